// sim_select_steps.cpp -- CPU simulation of the lock-step windowed quick-select used by stack_column_kernel
// (nl_column.cuh: qselect) on the synthetic benchmark columns: counts, per clipping pass, the steps every lane
// (pixel) needs and the steps a warp of 32 lanes executes (the maximum over its lanes), for a window width W.
// This is where the lane-utilisation figures of profiles/r01_summary.md and DESIGN.md come from.
//   g++ -O2 -ffp-contract=off -o sim tools/sim_select_steps.cpp oracle/libnl_oracle.so -Wl,-rpath,$PWD/oracle
//   ./sim 4        # window width (1 = one sample per pointer and step)
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cmath>
#include <algorithm>
extern "C" float nlo_synth_sample(uint32_t p, uint32_t k, uint32_t seed);
static int W=1;
struct Stat { long steps=0, r1=0; };
static float qselect_count(float *a, int n, int k, Stat &st) {
    int left=0,right=n-1; int round=0;
    while (left<right) {
        float pivot=a[(left+right)>>1];
        int l=left,r=right;
        for(;;){
            // one step: if both at stop -> swap/close; else each non-stopped side advances up to W
            bool sl=a[l]>=pivot, sr=a[r]<=pivot;
            st.steps++; if(round==0) st.r1++;
            if (sl&&sr){ if(l<r){ std::swap(a[l],a[r]); l++; r--; 
                    // after swap continue scanning W-1 more in same step
                    for(int j=0;j<W-1 && !(a[l]>=pivot);j++) l++;
                    for(int j=0;j<W-1 && !(a[r]<=pivot);j++) r--;
                 } else break; }
            else { for(int j=0;j<W && !(a[l]>=pivot);j++) l++;  for(int j=0;j<W && !(a[r]<=pivot);j++) r--; }
        }
        round++; int off=r-left+1;
        if (k<=off) right=r; else { left=r+1; k-=off; }
    }
    return a[left];
}
int main(int argc,char**argv){
    W=atoi(argv[1]);
    const int N=256; const int P=32*1000;
    std::vector<float> col(N+32);
    double tot_lane[5]={0}, tot_warp[5]={0}, tot_r1[5]={0}, tot_wr1[5]={0}; long wr1[5]; long npass[5]={0}; long wmax[5];
    for (int p=0;p<P;p++){
        if (p%32==0) for(int i=0;i<5;i++) {wmax[i]=0; wr1[i]=0;}
        int cur=0;
        for (int k=0;k<N;k++){ float v=nlo_synth_sample(p,k,12345); if (v==v) col[cur++]=v; }
        for (int pass=0; pass<5; pass++){
            Stat st; int kk=(cur>>1)+1;
            float up=qselect_count(col.data(),cur,kk,st);
            float med=up;
            if(!(cur&1)){ float lo=col[0]; for(int i=1;i<kk-1;i++) lo=std::max(lo,col[i]); med=0.5f*(lo+up);}            
            float s=0; for(int i=0;i<cur;i++) s+=col[i]; float m=s/cur; float v=0; for(int i=0;i<cur;i++){float d=col[i]-m; v+=d*d;} v/=cur; float sd=sqrtf(v);
            float lo=med-2.75f*sd, hi=med+2.75f*sd; int before=cur;
            for(int j=0;j<cur;){ if(col[j]<lo||col[j]>hi){ cur--; col[j]=col[cur]; } else j++; }
            tot_lane[pass]+=st.steps; tot_r1[pass]+=st.r1; wr1[pass]=std::max(wr1[pass],st.r1); npass[pass]++;
            wmax[pass]=std::max(wmax[pass],st.steps);
            if (cur==before||cur<=1) break;
        }
        if (p%32==31) for(int i=0;i<5;i++) {tot_warp[i]+=wmax[i]; tot_wr1[i]+=wr1[i];}
    }
    double tl=0,tw=0;
    for(int i=0;i<5;i++) if(npass[i]) { printf("W=%d pass %d: frac %.2f steps/lane %.0f (round1 %.0f) warp-max %.0f (round1 warp-max %.0f)\n",W,i,npass[i]/(double)P,tot_lane[i]/npass[i],tot_r1[i]/npass[i],tot_warp[i]/(P/32),tot_wr1[i]/(P/32)); tl+=tot_lane[i]/P; tw+=tot_warp[i]/(P/32);}
    printf("W=%d total useful steps/pixel %.0f, warp steps/tile %.0f\n",W,tl,tw);
}
