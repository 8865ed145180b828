// knn_model.cpp — CPU model of knn_warp_kernel's level selection (development tool, not product, not a test).
// Counts the work the warp-per-query kernel does on a given cloud so heuristics can be tuned without GPU time.
// usage: knn_model cloud.f32 (N x 4 floats) [variant]
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
typedef unsigned long long u64;
static int MB = 21;
static u64 expand21(unsigned v){u64 x=v&0x1FFFFFull;x=(x|x<<32)&0x1F00000000FFFFull;x=(x|x<<16)&0x1F0000FF0000FFull;x=(x|x<<8)&0x100F00F00F00F00Full;x=(x|x<<4)&0x10C30C30C30C30C3ull;x=(x|x<<2)&0x1249249249249249ull;return x;}
static u64 morton3(unsigned x,unsigned y,unsigned z){return expand21(x)|(expand21(y)<<1)|(expand21(z)<<2);}
static unsigned compact21(u64 x){x&=0x1249249249249249ull;x=(x^(x>>2))&0x10C30C30C30C30C3ull;x=(x^(x>>4))&0x100F00F00F00F00Full;x=(x^(x>>8))&0x1F0000FF0000FFull;x=(x^(x>>16))&0x1F00000000FFFFull;x=(x^(x>>32))&0x1FFFFFull;return (unsigned)x;}
struct P{float x,y,z;int id;};
int main(int argc,char**argv){
  FILE*f=fopen(argv[1],"rb");fseek(f,0,SEEK_END);long sz=ftell(f);fseek(f,0,SEEK_SET);int n=sz/16;std::vector<float>raw(n*4);fread(raw.data(),4,n*4,f);fclose(f);
  if(getenv("MB"))MB=atoi(getenv("MB")); int variant=argc>2?atoi(argv[2]):0; int K=20; int CAP=argc>3?atoi(argv[3]):1024; int MINTOT=argc>4?atoi(argv[4]):0; int QPW=argc>5?atoi(argv[5]):4;
  double mn[3]={1e30,1e30,1e30},mx[3]={-1e30,-1e30,-1e30};
  for(int i=0;i<n;i++)for(int d=0;d<3;d++){mn[d]=std::min(mn[d],(double)raw[i*4+d]);mx[d]=std::max(mx[d],(double)raw[i*4+d]);}
  double ext=0;for(int d=0;d<3;d++)ext=std::max(ext,mx[d]-mn[d]);ext=std::max(ext,1e-6)*(1+1e-6);double cell=ext/(double)(1<<MB),scale=1.0/cell;
  std::vector<std::pair<u64,int>> ord(n);
  for(int i=0;i<n;i++){unsigned q[3];for(int d=0;d<3;d++)q[d]=(unsigned)std::min(std::max(std::floor(((double)raw[i*4+d]-mn[d])*scale),0.0),(double)((1<<MB)-1));ord[i]={morton3(q[0],q[1],q[2]),i};}
  std::sort(ord.begin(),ord.end());
  std::vector<u64>codes(n);std::vector<P>pts(n);
  for(int j=0;j<n;j++){codes[j]=ord[j].first;int i=ord[j].second;pts[j]={raw[i*4],raw[i*4+1],raw[i*4+2],i};}
  auto lb=[&](u64 key){return (int)(std::lower_bound(codes.begin(),codes.end(),key)-codes.begin());};
  // counters
  double n_first=0,n_append=0,n_setup=0,n_batches=0,n_sorts=0,n_inserts=0,n_scans=0,n_cand=0,n_steer=0; long hist[12]={0};
  std::vector<u64> keys; keys.reserve(1<<16);
  int prevL=-1,pcx=0,pcy=0,pcz=0; int lo[27],cnt[27],total=0;
  for(int j=0;j<n;j++){
    if(j%QPW==0)prevL=-1;
    P q=pts[j];u64 cq=codes[j];unsigned ix=compact21(cq),iy=compact21(cq>>1),iz=compact21(cq>>2);
    int L;bool reuse=false;
    if(prevL>=0&&(int)(ix>>(MB-prevL))==pcx&&(int)(iy>>(MB-prevL))==pcy&&(int)(iz>>(MB-prevL))==pcz){L=prevL;reuse=true;}
    else{
      int W=4;
      u64 xl=cq^codes[std::max(j-W,0)],xr=cq^codes[std::min(j+W,n-1)];
      int bl=(j-W>=0)?(xl?(__builtin_clzll(xl)-(64-3*MB)):3*MB):0,br=(j+W<n)?(xr?(__builtin_clzll(xr)-(64-3*MB)):3*MB):0;
      L=std::min(std::max(bl,br)/3,MB);
      if(variant&1){ // start one level coarser than the window guess
        L=std::max(L-1,0);
      }
    }
    bool no_finer=reuse; int scans=0;
    for(;;){
      int sh=MB-L;int cx=ix>>sh,cy=iy>>sh,cz=iz>>sh;int ncell=1<<L;
      if(!reuse){
        n_setup++;
        total=0;
        for(int o=0;o<27;o++){int nx=cx+o%3-1,ny=cy+(o/3)%3-1,nz=cz+o/9-1;lo[o]=0;cnt[o]=0;
          if(nx>=0&&ny>=0&&nz>=0&&nx<ncell&&ny<ncell&&nz<ncell){u64 pre=morton3(nx,ny,nz);lo[o]=lb(pre<<(3*sh));int hi=(L==0)?n:lb((pre+1ull)<<(3*sh));cnt[o]=hi-lo[o];}
          total+=cnt[o];}
        if(!no_finer&&total>CAP&&L<MB){L++;n_steer++;continue;}
        if(total<K&&L>0){no_finer=true;L--;n_steer++;continue;}
        if((variant&2)&&total<MINTOT&&L>0&&!(scans>0)){ // steer coarser when the block is unlikely to pass the exactness test
          no_finer=true;L--;n_steer++;continue;}
      }
      reuse=false;prevL=L;pcx=cx;pcy=cy;pcz=cz;
      scans++;n_scans++;n_cand+=total;
      // gather candidates in kernel order (centre first)
      keys.clear();
      static const int order_[27]={13,0,1,2,3,4,5,6,7,8,9,10,11,12,14,15,16,17,18,19,20,21,22,23,24,25,26};
      for(int t=0;t<27;t++){int o=order_[t];for(int c=lo[o];c<lo[o]+cnt[o];c++){float dx=q.x-pts[c].x,dy=q.y-pts[c].y,dz=q.z-pts[c].z;float d=dx*dx;d=d+dy*dy;d=d+dz*dz;uint32_t db;memcpy(&db,&d,4);keys.push_back(((u64)db<<32)|(unsigned)pts[c].id);}}
      std::vector<u64> top(32,0x7F7FFFFF7FFFFFFFull);
      int THR=(variant&8)?10:6;
      if(variant&32){
        // window-first: the 32 sorted-order neighbours of the query form the first batch (one contiguous load, no cell mapping)
        int w0=std::max(0,std::min(j-16,n-32)), w1=std::min(n,w0+32);
        std::vector<u64> wk;
        for(int c=w0;c<w1;c++){float dx=q.x-pts[c].x,dy=q.y-pts[c].y,dz=q.z-pts[c].z;float d=dx*dx;d=d+dy*dy;d=d+dz*dz;uint32_t db;memcpy(&db,&d,4);wk.push_back(((u64)db<<32)|(unsigned)pts[c].id);}
        n_first++; n_batches+=0.6;   // cheaper batch: no owner search
        for(u64 k2:wk)top.push_back(k2);std::sort(top.begin(),top.end());top.resize(32);
        // block scan, skipping candidates whose sorted position lies inside the window
        std::vector<u64> keys2; int t=0;
        static const int order2[27]={13,0,1,2,3,4,5,6,7,8,9,10,11,12,14,15,16,17,18,19,20,21,22,23,24,25,26};
        for(size_t base=0;base<keys.size();base+=32){
          n_batches++;
          u64 kth=top[K-1];int pass=0;std::vector<u64>pk;
          for(size_t g=base;g<std::min(base+32,keys.size());g++){
            // recover sorted position of candidate g
            int rem=(int)g,pos=-1;for(int tt=0;tt<27;tt++){int o=order2[tt];if(rem<cnt[o]){pos=lo[o]+rem;break;}rem-=cnt[o];}
            if(pos>=w0&&pos<w1)continue;
            if(keys[g]<kth){pass++;pk.push_back(keys[g]);}
          }
          if(!pass)continue;
          if(pass>THR)n_sorts++;else n_inserts+=pass;
          for(u64 k2:pk)top.push_back(k2);std::sort(top.begin(),top.end());top.resize(32);
        }
        (void)t;
      } else
      if(!(variant&4)){
      for(size_t base=0;base<keys.size();base+=32){
        n_batches++;
        u64 kth=top[K-1];int pass=0;std::vector<u64>pk;
        for(size_t g=base;g<std::min(base+32,keys.size());g++)if(keys[g]<kth){pass++;pk.push_back(keys[g]);}
        if(!pass)continue;
        if(pass>THR){ if(base==0) n_first++; else n_sorts++;}else n_inserts+=pass;
        for(u64 k2:pk)top.push_back(k2);std::sort(top.begin(),top.end());top.resize(32);
      }
      } else {
        std::vector<u64> pend; u64 kth=top[K-1];
        auto flush=[&](bool first){ if(pend.empty())return; if((int)pend.size()>THR){ if(first) n_first++; else n_sorts++; } else n_inserts+=pend.size();
          for(u64 k2:pend)top.push_back(k2);std::sort(top.begin(),top.end());top.resize(32);pend.clear();kth=top[K-1]; };
        for(size_t base=0;base<keys.size();base+=32){
          n_batches++;
          std::vector<u64>pk;
          for(size_t g=base;g<std::min(base+32,keys.size());g++)if(keys[g]<kth)pk.push_back(keys[g]);
          if(pk.empty())continue;
          n_append++;
          if(pend.size()+pk.size()>32){ flush(false); std::vector<u64> pk2; for(u64 k2:pk) if(k2<kth) pk2.push_back(k2); pk=pk2; }
          for(u64 k2:pk)pend.push_back(k2);
          if(base==0||pend.size()==32) flush(base==0);
        }
        flush(false);
      }
      if(L==0)break;
      double sL=cell*(double)(1<<sh);double rel[3]={(double)q.x-mn[0],(double)q.y-mn[1],(double)q.z-mn[2]};double rs=1e300;int cc[3]={cx,cy,cz};
      for(int dd=0;dd<3;dd++){if(cc[dd]-1>=0)rs=std::min(rs,rel[dd]-(double)(cc[dd]-1)*sL);if(cc[dd]+1<ncell)rs=std::min(rs,(double)(cc[dd]+2)*sL-rel[dd]);}
      u64 kth=top[K-1];float kd;uint32_t kb=(uint32_t)(kth>>32);memcpy(&kd,&kb,4);
      if(kth!=0x7F7FFFFF7FFFFFFFull&&(double)kd<=rs*rs*(1-1e-6))break;
      no_finer=true;L--;
    }
    hist[std::min(scans,11)]++;
  }
  double N=n;
  double cost=n_setup/N*160+n_batches/N*45+n_sorts/N*170+n_first/N*((variant&16)?125:170)+n_append/N*((variant&4)?12:0)+n_inserts/N*14+60;
  printf("variant %d CAP %d MINTOT %d QPW %d | per query: setups %.2f steer %.2f scans %.2f cand %.0f batches %.1f first %.2f append %.2f sorts %.2f inserts %.1f | model cost %.0f instr\n",variant,CAP,MINTOT,QPW,n_setup/N,n_steer/N,n_scans/N,n_cand/N,n_batches/N,n_first/N,n_append/N,n_sorts/N,n_inserts/N,cost);
  printf("scans histogram:");for(int i=0;i<12;i++)printf(" %ld",hist[i]);printf("\n");
}
