// sim_block_partition.cpp -- CPU simulation of a BLOCK-PARTITION quick-select (masks of stops per B-slot block, k-th left
// stop paired with k-th right stop) run in lock step over 32 lanes on the benchmark columns: checks that it leaves the
// oracle's permutation and estimates its warp instructions per tile (lever (i)).  Result: profiles/r02_select_levers.md
//   g++ -O2 -ffp-contract=off -o sim tools/sim_block_partition.cpp oracle/libnl_oracle.so -Wl,-rpath,$PWD/oracle; ./sim 32
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cmath>
#include <algorithm>
#include <cstring>
extern "C" float nlo_synth_sample(uint32_t p, uint32_t k, uint32_t seed);
extern "C" float nlo_qselect_f32(float*, int, int);
static int B = 16;
struct Lane {
    std::vector<float> a; int n; int left, right, k; float pivot; int l, r; bool active;
    long iters=0, swaps=0;
    void begin(int n_, int k_) { n=n_; left=0; right=n-1; k=k_; active = n>1; if(n>0) pivot=a[(left+right)>>1]; l=left; r=right; }
};
// one iteration for a lane; returns number of swaps done (pairing loop trips)
static int step(Lane &L) {
    if (!L.active) return 0;
    L.iters++;
    float *a = L.a.data();
    uint64_t ml=0, mr=0;
    for (int j=0;j<B;j++){ int s=L.l+j; if (s<=L.right && a[s]>=L.pivot) ml|=1ull<<j; }
    for (int j=0;j<B;j++){ int s=L.r-j; if (s>=L.left && a[s]<=L.pivot) mr|=1ull<<j; }
    int sl_last=-1000000, sr_last=1000000; int sw=0;
    while (ml && mr) {
        int pl=L.l+__builtin_ctzll(ml), pr=L.r-__builtin_ctzll(mr);
        if (pl>=pr) break;
        std::swap(a[pl],a[pr]); ml&=ml-1; mr&=mr-1; sl_last=pl; sr_last=pr; sw++;
    }
    L.swaps+=sw;
    bool lk, rk; int Lt, Rt;
    if (ml) { lk=true; Lt=std::min(L.l+__builtin_ctzll(ml), sr_last); }
    else if (sr_last<=L.l+B-1) { lk=true; Lt=sr_last; } else { lk=false; Lt=L.l+B; }
    if (mr) { rk=true; Rt=std::max(L.r-__builtin_ctzll(mr), sl_last); }
    else if (sl_last>=L.r-B+1) { rk=true; Rt=sl_last; } else { rk=false; Rt=L.r-B; }
    if (lk && rk) {
        // crossed
        int index=Rt; int off=index-L.left+1;
        if (L.k<=off) L.right=index; else { L.left=index+1; L.k-=off; }
        L.active = L.left<L.right;
        L.pivot=a[(L.left+L.right)>>1]; L.l=L.left; L.r=L.right;
    } else { L.l=Lt; L.r=Rt; }
    return sw;
}
int main(int argc,char**argv){
    B=atoi(argv[1]); const int N=256; const int TILES=400;
    const double C_BLOCK = 6.0*B + 40, C_SWAP=14;
    double tot_instr[6]={0}, tot_iters[6]={0}, tot_sw[6]={0}, lane_iters[6]={0}, lane_sw[6]={0}; long tiles_pass[6]={0}; long lanes_pass[6]={0};
    long mism=0;
    for (int t=0;t<TILES;t++){
        Lane L[32]; int cur[32]; bool done[32];
        for (int i=0;i<32;i++){ L[i].a.assign(N+64,0); cur[i]=0; done[i]=false; for(int k=0;k<N;k++){ float v=nlo_synth_sample(t*32+i,k,12345); if(v==v) L[i].a[cur[i]++]=v; } }
        for (int pass=0;pass<6;pass++){
            bool any=false; for(int i=0;i<32;i++) any|=!done[i]; if(!any) break;
            tiles_pass[pass]++;
            std::vector<std::vector<float>> ref(32);
            for (int i=0;i<32;i++){ int m=done[i]?0:cur[i]; ref[i].assign(L[i].a.begin(), L[i].a.begin()+m); if(m>0) nlo_qselect_f32(ref[i].data(), m, (m>>1)+1);
                L[i].begin(m,(m>>1)+1); L[i].iters=0; L[i].swaps=0; if(!done[i]) lanes_pass[pass]++; }
            long witers=0, wsw=0;
            for(;;){ bool act=false; int mx=0; for(int i=0;i<32;i++){ if(L[i].active){act=true; int s=step(L[i]); mx=std::max(mx,s);} } if(!act) break; witers++; wsw+=mx; }
            for (int i=0;i<32;i++){ int m=done[i]?0:cur[i]; if (m>0 && memcmp(ref[i].data(), L[i].a.data(), m*4)!=0) mism++; lane_iters[pass]+=L[i].iters; lane_sw[pass]+=L[i].swaps; }
            tot_iters[pass]+=witers; tot_sw[pass]+=wsw; tot_instr[pass]+=witers*C_BLOCK+wsw*C_SWAP;
            // the rest of the pass: median/mean/sd/clip
            for (int i=0;i<32;i++){ if(done[i]) continue; float *a=L[i].a.data(); int m=cur[i]; int kk=(m>>1)+1; float up=a[kk-1]; // after select a[left] where left==k-1
                float med=up; if(!(m&1)){ float lo=a[0]; for(int j=1;j<kk-1;j++) if(a[j]>lo) lo=a[j]; med=0.5f*(lo+up);} 
                float s=0; for(int j=0;j<m;j++) s+=a[j]; float mean=s/m; float v=0; for(int j=0;j<m;j++){float d=a[j]-mean; v+=d*d;} v/=m; float sd=sqrtf(v);
                float lo=med-2.75f*sd, hi=med+2.75f*sd; int before=m;
                for(int j=0;j<m;){ if(a[j]<lo||a[j]>hi){ m--; a[j]=a[m]; } else j++; }
                cur[i]=m; if(m==before||m<=1) done[i]=true; }
        }
    }
    printf("B=%d mismatches %ld\n",B,mism);
    double total=0;
    for(int p=0;p<6;p++) if(tiles_pass[p]) { printf("pass %d: tiles %.2f lanes/tile %.1f | warp iters %.1f (lane mean %.1f) pair trips %.1f (lane mean swaps %.1f) -> est instr/tile %.0f\n",p,tiles_pass[p]/(double)TILES, lanes_pass[p]/(double)tiles_pass[p], tot_iters[p]/tiles_pass[p], lane_iters[p]/lanes_pass[p], tot_sw[p]/tiles_pass[p], lane_sw[p]/lanes_pass[p], tot_instr[p]/tiles_pass[p]); total+=tot_instr[p]/TILES; }
    printf("B=%d est select instr per tile (all passes, no deferral) %.0f\n",B,total);
}
