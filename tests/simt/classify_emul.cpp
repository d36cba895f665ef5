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
                                );
        else
          wb_classify_kernel<2>(sx,sy,sz,n,nChunks,bounds.data(),levelOff.data(),levelCnt.data(),nLevels,winner.data(),hyp,
                                maxSlope,thickness,clsIn.data(),perm.data(),0u,0xffffffffu,labelSorted,counters,
                                wedge.data(),pending.data()
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
