// sim_phase_warps.cpp -- what would phase-specialised warps buy?  (result in profiles/r02_select_levers.md)
//   g++ -O2 -ffp-contract=off -o sim tools/sim_phase_warps.cpp oracle/libnl_oracle.so -Wl,-rpath,$PWD/oracle; ./sim <select warps> <regular warps> <tiles>
// Per SM: 7 slabs x 32 banks of columns; NS select warps and NR
// regular warps; lane i of a warp only ever works on a column of bank i (any slab).  Every warp runs at a fixed rate
// (3.7 cycles per instruction: the measured per-warp rate).  Columns: the benchmark data, steps per pass from the W=4
// window model, 36 instructions per select step, REG instructions per regular pass (mean/var/clip), INIT for staging.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cmath>
#include <algorithm>
#include <queue>
extern "C" float nlo_synth_sample(uint32_t p, uint32_t k, uint32_t seed);
static const int W=4;
static long qsteps(float *a, int n, int k) {
    long steps=0; int left=0,right=n-1;
    while (left<right) {
        float pivot=a[(left+right)>>1]; int l=left,r=right;
        for(;;){ bool sl=a[l]>=pivot, sr=a[r]<=pivot; steps++;
            if (sl&&sr){ if(l<r){ std::swap(a[l],a[r]); l++; r--; for(int j=0;j<W-1 && !(a[l]>=pivot);j++) l++; for(int j=0;j<W-1 && !(a[r]<=pivot);j++) r--; } else break; }
            else { for(int j=0;j<W && !(a[l]>=pivot);j++) l++;  for(int j=0;j<W && !(a[r]<=pivot);j++) r--; } }
        int off=r-left+1; if (k<=off) right=r; else { left=r+1; k-=off; }
    }
    return steps;
}
struct Col { std::vector<int> steps; };   // select steps per pass (at most 3 passes here: later ones go to the pool)
int main(int argc,char**argv){
    int NS=atoi(argv[1]), NR=atoi(argv[2]); const int N=256, TILES=atoi(argv[3]); const int MAXP=3;
    const double STEP=36, REG=2700, INIT=1500, CLOSE=25, CLAIM=25;
    std::vector<Col> cols((size_t)TILES*32);
    std::vector<float> col(N+32);
    for (int p=0;p<TILES*32;p++){ int cur=0; for(int k=0;k<N;k++){ float v=nlo_synth_sample(p,k,12345); if(v==v) col[cur++]=v; }
        for(int pass=0;pass<MAXP;pass++){ long st=qsteps(col.data(),cur,(cur>>1)+1); cols[p].steps.push_back((int)st);
            int kk=(cur>>1)+1; float up=col[kk-1], med=up; if(!(cur&1)){ float lo=col[0]; for(int i=1;i<kk-1;i++) lo=std::max(lo,col[i]); med=0.5f*(lo+up);} 
            float s=0; for(int i=0;i<cur;i++) s+=col[i]; float m=s/cur; float v=0; for(int i=0;i<cur;i++){float d=col[i]-m; v+=d*d;} v/=cur; float sd=sqrtf(v);
            float lo=med-2.75f*sd, hi=med+2.75f*sd; int before=cur; for(int j=0;j<cur;){ if(col[j]<lo||col[j]>hi){ cur--; col[j]=col[cur]; } else j++; }
            if (cur==before||cur<=1) break; } }
    // ---- baseline: a warp owns a slab; per pass cost = max over lanes of steps * STEP + closes + REG
    double base=0; for(int t=0;t<TILES;t++){ double c=INIT; for(int pass=0;pass<MAXP;pass++){ int mx=0; bool any=false; for(int i=0;i<32;i++){ auto&s=cols[t*32+i].steps; if((int)s.size()>pass){any=true; mx=std::max(mx,s[pass]);} } if(!any) break; c+=mx*STEP+ (mx/6.0)*CLOSE + REG; } base+=c; }
    double base_per_sm = base/7.0;   // 7 warps in parallel
    // ---- new: event simulation
    struct Slab { int tile=-1; int done=32; };
    std::vector<Slab> slabs(7);
    // column state per slab/bank: 0 empty/done, 1 need_init, 2 need_select, 3 selecting, 4 need_regular, 5 in regular
    int state[7][32]; int pass_[7][32]; double rem[7][32];
    for(int s=0;s<7;s++) for(int b=0;b<32;b++){ state[s][b]=0; pass_[s][b]=0; rem[s][b]=0; }
    int next_tile=0; int retired=0;
    struct Warp { int role; double t; int job_s[32]; };
    std::vector<Warp> warps(NS+NR);
    for(int w=0;w<NS+NR;w++){ warps[w].role = w<NS?0:1; warps[w].t=0; for(int i=0;i<32;i++) warps[w].job_s[i]=-1; }
    double busy_sel=0, tot_sel=0, busy_reg=0, tot_reg=0;
    long done_tiles=0; double tend=0;
    while (retired<7) {
        // pick the warp with the smallest clock
        int w=0; for(int i=1;i<NS+NR;i++) if (warps[i].t<warps[w].t) w=i;
        Warp &wp=warps[w];
        if (wp.role==0) {
            // claim
            bool claimed=false; int active=0;
            for(int b=0;b<32;b++){ if (wp.job_s[b]<0){ for(int s=0;s<7;s++) if(state[s][b]==2){ state[s][b]=3; wp.job_s[b]=s; rem[s][b]=cols[slabs[s].tile*32+b].steps[pass_[s][b]]; claimed=true; break; } } if(wp.job_s[b]>=0) active++; }
            if (!active){ wp.t+=200; continue; }
            double cost=6*STEP+CLOSE+(claimed?CLAIM:0);
            busy_sel+=active*6; tot_sel+=32*6;
            for(int b=0;b<32;b++){ int s=wp.job_s[b]; if(s<0) continue; rem[s][b]-=6; if(rem[s][b]<=0){ state[s][b]=4; wp.job_s[b]=-1; } }
            wp.t+=cost;
        } else {
            // regular: claim need_init / need_regular columns, one per lane
            int jobs[32]; int active=0; bool anyinit=false, anyreg=false;
            for(int b=0;b<32;b++){ jobs[b]=-1; for(int s=0;s<7;s++) if(state[s][b]==1||state[s][b]==4){ jobs[b]=s; if(state[s][b]==1) anyinit=true; else anyreg=true; break; } if(jobs[b]>=0) active++; }
            if (!active) {
                // refill a finished slab
                bool did=false;
                for(int s=0;s<7;s++) if(slabs[s].done==32 && slabs[s].tile!=-2){ if(next_tile<TILES){ if(slabs[s].tile>=0) done_tiles++; slabs[s].tile=next_tile++; slabs[s].done=0; for(int b=0;b<32;b++){ state[s][b]=1; pass_[s][b]=0; } wp.t+=300; did=true; } else { if(slabs[s].tile>=0) done_tiles++; slabs[s].tile=-2; retired++; tend=std::max(tend,wp.t); } break; }
                if(!did) wp.t+=200;
                continue;
            }
            double cost=CLAIM+(anyinit?INIT:0)+(anyreg?REG:0);
            busy_reg+=active; tot_reg+=32;
            for(int b=0;b<32;b++){ int s=jobs[b]; if(s<0) continue; if(state[s][b]==1){ state[s][b]=2; } else { pass_[s][b]++; auto&st=cols[slabs[s].tile*32+b].steps; if(pass_[s][b]<(int)st.size()) state[s][b]=2; else { state[s][b]=0; slabs[s].done++; } } }
            wp.t+=cost;
        }
    }
    printf("NS=%d NR=%d: baseline %.0f instr-time per SM (7 warps), new %.0f  -> speedup %.2fx ; select lane util %.2f regular lane util %.2f\n", NS,NR, base_per_sm, tend, base_per_sm/tend, busy_sel/tot_sel, busy_reg/tot_reg);
}
