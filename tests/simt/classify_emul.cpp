// classify_emul.cpp — runs the classify kernels of wolkenbase_b200/csrc/wb_kernels.cuh (the very source nvcc
// compiles for sm_100a) on the CPU, one emulated warp at a time.  TEST INFRASTRUCTURE ONLY: it lets the CPU
// test suite check the kernel's traversal logic (and variants of it behind WB_CL_* macros) against the oracle's
// labels before a GPU is at hand.  The GPU tests remain the parity proof for the compiled kernel.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#include <cstdio>
#include "cuda_runtime.h"
#include "../../wolkenbase_b200/csrc/wb_kernels.cuh"

extern "C" int simt_set_tables(const double *tanTable,const double *cosTable,const double *sinTable)
// angle.cpp:305-320 tables, as the library uploads them in wb_create (511/512/512 entries)
{
  memset(g_tanTable,0,sizeof(g_tanTable));
  memcpy(g_tanTable,tanTable,511*sizeof(double));
  memcpy(g_cosTable,cosTable,512*sizeof(double));
  memcpy(g_sinTable,sinTable,512*sizeof(double));
  return 0;
}

extern "C" void simt_site_counts(unsigned long long *out /* 4096, indexed by source line of wb_kernels.cuh mod 4096 */,int clear)
{
  memcpy(out,simt::site_counts(),4096*sizeof(unsigned long long));
  if (clear)
    memset(simt::site_counts(),0,4096*sizeof(unsigned long long));
}

extern "C" void simt_emu_counts(unsigned long long *out /* 16 */,int clear)
{
  memcpy(out,simt::emu_counts(),16*sizeof(unsigned long long));
  if (clear)
    memset(simt::emu_counts(),0,16*sizeof(unsigned long long));
}

#if WB_CL_PLANES
static std::vector<WbPlane> g_planes;
// Prototype of the plane builder (plain loops; the CUDA build kernels are round-2 work): per chunk a trimmed
// least-squares slope with the intercept lowered until every point is on or above the plane, per node the mean of
// the children's slopes lowered under every child plane over that child's box; the flat plane z = zmin stays
// whenever it is the higher one at the box centre.
static void fitSlopes(const std::vector<double> &u,const std::vector<double> &v,const std::vector<double> &z,
                      const std::vector<char> &use,double &b,double &c)
{
  double n=0,su=0,sv=0,sz=0;
  for (size_t i=0;i<u.size();i++) if (use[i]) { n++; su+=u[i]; sv+=v[i]; sz+=z[i]; }
  b=c=0;
  if (n<3) return;
  su/=n; sv/=n; sz/=n;
  double uu=0,uv=0,vv=0,uz=0,vz=0;
  for (size_t i=0;i<u.size();i++) if (use[i])
  {
    double du=u[i]-su,dv=v[i]-sv,dz=z[i]-sz;
    uu+=du*du; uv+=du*dv; vv+=dv*dv; uz+=du*dz; vz+=dv*dz;
  }
  double det=uu*vv-uv*uv;
  if (!(det>1e-9*(uu*vv+1e-30))) return;
  b=(uz*vv-vz*uv)/det;
  c=(vz*uu-uz*uv)/det;
  if (!(fabs(b)<4 && fabs(c)<4)) b=c=0;          // walls: no useful slope
}

static void buildPlanes(const double *sx,const double *sy,const double *sz,uint64_t n,const std::vector<WbBound> &bounds,
                        const std::vector<uint32_t> &levelOff,const std::vector<uint32_t> &levelCnt,int nLevels,
                        std::vector<WbPlane> &planes)
{
  planes.assign(bounds.size(),WbPlane{0,0,0});
  const uint32_t nChunks=levelCnt[0];
  std::vector<double> u,v,z,r;
  std::vector<char> use;
  for (uint32_t k=0;k<nChunks;k++)
  {
    const WbBound &bd=bounds[k];
    const double xc=0.5*(bd.xmin+bd.xmax),yc=0.5*(bd.ymin+bd.ymax);
    u.clear(); v.clear(); z.clear();
    for (uint64_t j=(uint64_t)k*32;j<n && j<(uint64_t)k*32+32;j++) { u.push_back(sx[j]-xc); v.push_back(sy[j]-yc); z.push_back(sz[j]); }
    use.assign(u.size(),1);
    double b=0,c=0;
    for (int it=0;it<3;it++)
    {
      fitSlopes(u,v,z,use,b,c);
      r.resize(u.size());
      for (size_t i=0;i<u.size();i++) r[i]=z[i]-b*u[i]-c*v[i];
      std::vector<double> sr(r);
      std::nth_element(sr.begin(),sr.begin()+sr.size()/2,sr.end());
      const double med=sr[sr.size()/2];
      for (size_t i=0;i<u.size();i++) use[i]=r[i]<=med;          // the lower half: ground, not what grows on it
    }
    double a=INFINITY;
    for (size_t i=0;i<u.size();i++) a=fmin(a,z[i]-b*u[i]-c*v[i]);
    a-=1e-9;
    if (a>bd.zmin) planes[k]=WbPlane{a,(float)b,(float)c};
    else planes[k]=WbPlane{bd.zmin,0,0};
    // the stored slopes are floats: lower a until the plane with the ROUNDED slopes is under every point
    if (planes[k].b!=0 || planes[k].c!=0)
    {
      double aa=INFINITY;
      for (size_t i=0;i<u.size();i++) aa=fmin(aa,z[i]-(double)planes[k].b*u[i]-(double)planes[k].c*v[i]);
      planes[k].a=aa-1e-9;
    }
  }
  for (int l=1;l<nLevels;l++)
    for (uint32_t k=0;k<levelCnt[l];k++)
    {
      const WbBound &bd=bounds[levelOff[l]+k];
      const double xc=0.5*(bd.xmin+bd.xmax),yc=0.5*(bd.ymin+bd.ymax);
      double sb=0,sc=0,cnt=0;
      for (uint32_t j=k*32;j<levelCnt[l-1] && j<k*32+32;j++) { sb+=planes[levelOff[l-1]+j].b; sc+=planes[levelOff[l-1]+j].c; cnt++; }
      const float b=(float)(sb/cnt),c=(float)(sc/cnt);
      double a=INFINITY;
      for (uint32_t j=k*32;j<levelCnt[l-1] && j<k*32+32;j++)
      {
        const WbBound &cb=bounds[levelOff[l-1]+j];
        const WbPlane &cp=planes[levelOff[l-1]+j];
        const double cxc=0.5*(cb.xmin+cb.xmax),cyc=0.5*(cb.ymin+cb.ymax);
        for (int corner=0;corner<4;corner++)
        {
          const double x=(corner&1)?cb.xmax:cb.xmin,y=(corner&2)?cb.ymax:cb.ymin;
          const double child=cp.a+(double)cp.b*(x-cxc)+(double)cp.c*(y-cyc);
          a=fmin(a,child-(double)b*(x-xc)-(double)c*(y-yc));
        }
      }
      a-=1e-9;
      planes[levelOff[l]+k]=(a>bd.zmin)?WbPlane{a,b,c}:WbPlane{bd.zmin,0,0};
    }
  // self-check: every point on or above the plane of its chunk and of every ancestor
  for (uint64_t j=0;j<n;j++)
  {
    uint32_t node=(uint32_t)(j/32);
    for (int l=0;l<nLevels;l++)
    {
      const WbBound &bd=bounds[levelOff[l]+node];
      const WbPlane &pl=planes[levelOff[l]+node];
      const double low=pl.a+(double)pl.b*(sx[j]-0.5*(bd.xmin+bd.xmax))+(double)pl.c*(sy[j]-0.5*(bd.ymin+bd.ymax));
      if (sz[j]<low-1e-7) { fprintf(stderr,"plane above point %llu at level %d by %g\n",(unsigned long long)j,l,low-sz[j]); abort(); }
      node/=32;
    }
  }
}
#endif

static bool g_reset=true;
extern "C" void simt_reset() { g_reset=true; }   // the caller's arrays changed: forget the cached hierarchy

template <typename F> static unsigned long long launch(unsigned nBlocks,unsigned blockThreads,F body)
// every block = blockThreads/32 independent warps (none of the kernels run here uses __syncthreads)
{
  unsigned long long collectives=0;
  for (unsigned b=0;b<nBlocks;b++)
    for (unsigned w=0;w<blockThreads/32;w++)
      collectives+=simt::run_warp(body,w*32,b,blockThreads,nBlocks);
  return collectives;
}

extern "C" int simt_classify(const double *sx,const double *sy,const double *sz,uint64_t n,
                             const double *hyp,          // per point: hyperboloidSize of its tile, NaN = in no tile
                             double maxSlope,double thickness,
                             uint32_t firstChunk,uint32_t endChunk,   // label only the queries of these chunks (all: 0,~0)
                             uint8_t *labelSorted,       // n bytes; untiled points get 0, chunks outside the range 254
                             unsigned long long *counters /* 24 */,unsigned long long *collectives)
// Chunk bounds and the 32-ary hierarchy (wb_chunk_bounds_kernel, wb_node_bounds_kernel), then
// wb_classify_kernel<1> and <2>, exactly as wb_build/wb_classify launch them (wolken_b200.cu).
{
  const uint32_t nChunks=(uint32_t)((n+31)/32);
  std::vector<uint32_t> levelOff,levelCnt;
  uint64_t total=0;
  uint32_t c=nChunks;
  while (true)
  {
    levelOff.push_back((uint32_t)total);
    levelCnt.push_back(c);
    total+=c;
    if (c<=32)
      break;
    c=(c+31)/32;
  }
  const int nLevels=(int)levelCnt.size();
  // The hierarchy is kept between calls on the same arrays (a caller sampling warps of a large cloud).  Small
  // clouds go through the emulated wb_chunk_bounds_kernel / wb_node_bounds_kernel; large ones through plain loops
  // that take the same minima and maxima.
  static std::vector<WbBound> bounds;
  static const double *boundsOf=nullptr;
  static uint64_t boundsN=0;
  levelOff.resize(16); levelCnt.resize(16);
  unsigned long long coll=0;
  if (g_reset || boundsOf!=sx || boundsN!=n)
  {
    bounds.assign(total,WbBound());
    if (n<=(1u<<20))
    {
      coll+=launch((nChunks*32+255)/256,256,[&]{ wb_chunk_bounds_kernel(sx,sy,sz,n,bounds.data(),nChunks); });
      for (int l=1;l<nLevels;l++)
        coll+=launch((levelCnt[l]*32+255)/256,256,[&]{ wb_node_bounds_kernel(bounds.data()+levelOff[l-1],levelCnt[l-1],
                                                                               bounds.data()+levelOff[l],levelCnt[l]); });
    }
    else
    {
      for (uint32_t k=0;k<nChunks;k++)
      {
        WbBound b={INFINITY,-INFINITY,INFINITY,-INFINITY,INFINITY};
        for (uint64_t j=(uint64_t)k*32;j<n && j<(uint64_t)k*32+32;j++)
        {
          b.xmin=fmin(b.xmin,sx[j]); b.xmax=fmax(b.xmax,sx[j]);
          b.ymin=fmin(b.ymin,sy[j]); b.ymax=fmax(b.ymax,sy[j]);
          b.zmin=fmin(b.zmin,sz[j]);
        }
        bounds[k]=b;
      }
      for (int l=1;l<nLevels;l++)
        for (uint32_t k=0;k<levelCnt[l];k++)
        {
          WbBound b={INFINITY,-INFINITY,INFINITY,-INFINITY,INFINITY};
          for (uint32_t j=k*32;j<levelCnt[l-1] && j<k*32+32;j++)
          {
            const WbBound &ch=bounds[levelOff[l-1]+j];
            b.xmin=fmin(b.xmin,ch.xmin); b.xmax=fmax(b.xmax,ch.xmax);
            b.ymin=fmin(b.ymin,ch.ymin); b.ymax=fmax(b.ymax,ch.ymax);
            b.zmin=fmin(b.zmin,ch.zmin);
          }
          bounds[levelOff[l]+k]=b;
        }
    }
    boundsOf=sx;
    boundsN=n;
#if WB_CL_PLANES
    buildPlanes(sx,sy,sz,n,bounds,levelOff,levelCnt,nLevels,g_planes);
#endif
  }
  // one "tile" per point: winner = own index (or none), tHyp = its hyperboloidSize
  static std::vector<uint32_t> winner,perm,wedge;
  static std::vector<uint8_t> clsIn,pending;
  static const double *hypOf=nullptr;
  if (g_reset || hypOf!=hyp || winner.size()!=n)
  {
    winner.resize(n); perm.resize(n); wedge.assign(n,0xffffffffu); clsIn.assign(n,0); pending.assign(nChunks,0);
    for (uint64_t i=0;i<n;i++)
    {
      winner[i]=std::isnan(hyp[i])?0xffffffffu:(uint32_t)i;
      perm[i]=(uint32_t)i;
    }
    hypOf=hyp;
  }
  g_reset=false;
  if (firstChunk==0 && endChunk>=nChunks)
    memset(labelSorted,254,n);
  if (endChunk>nChunks)
    endChunk=nChunks;
  // the kernels index chunks by blockIdx.x: run the requested range only (every other point still takes part
  // as a candidate)
  auto run=[&](int pass)
  {
    for (uint32_t b=firstChunk;b<endChunk;b++)
    {
      auto body=[&]
      {
        if (pass==1)
          wb_classify_kernel<1>(sx,sy,sz,n,nChunks,bounds.data(),levelOff.data(),levelCnt.data(),nLevels,winner.data(),hyp,
                                maxSlope,thickness,clsIn.data(),perm.data(),0u,0xffffffffu,labelSorted,counters,
                                wedge.data(),pending.data()
#if WB_CL_PLANES
                                ,g_planes.data()
#endif
                                );
        else
          wb_classify_kernel<2>(sx,sy,sz,n,nChunks,bounds.data(),levelOff.data(),levelCnt.data(),nLevels,winner.data(),hyp,
                                maxSlope,thickness,clsIn.data(),perm.data(),0u,0xffffffffu,labelSorted,counters,
                                wedge.data(),pending.data()
#if WB_CL_PLANES
                                ,g_planes.data()
#endif
                                );
      };
      coll+=simt::run_warp(body,0,b,WB_CL_WARPS*32,nChunks);
    }
  };
  static_assert(WB_CL_WARPS==1,"the emulator launches one warp per block");
  run(1);
  run(2);
  if (collectives)
    *collectives=coll;
  return 0;
}
