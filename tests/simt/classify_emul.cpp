// classify_emul.cpp — kernel-level harnesses over the SIMT emulator: the classify kernels of
// wolkenbase_b200/csrc/wb_kernels.cuh (the very source nvcc compiles for sm_100a) on their own, the tile phases + classify
// launched as wb_scan/wb_postscan/wb_classify launch them, the radix sort, the LAS decode.  TEST INFRASTRUCTURE ONLY: the
// CPU suite checks the kernels' logic (and variants behind WB_CL_* macros) against the oracle before a GPU is at
// hand, and tools/model_bench_scene.py counts their work on the real bench scene.  (The whole library, host code
// included, is the other target of the Makefile.)  The GPU tests remain the parity proof for the compiled kernels.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#include <cstdio>
#include "cuda_runtime.h"
#include "../../wolkenbase_b200/csrc/wb_kernels.cuh"
#define WB_NO_HOST_LAUNCH
#include "../../wolkenbase_b200/csrc/wb_sort.cuh"
#include "../../wolkenbase_b200/csrc/wb_host.h"

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
#if WB_CL_COMPACT2
    // pass 2 over the pending queries of the chunk range, gathered into full warps (wb_classify does the same with
    // wb_pending_flag_kernel, a scan and wb_pending_scatter_kernel)
    std::vector<uint32_t> pendingList;
    if (pass==2)
      for (uint64_t i=(uint64_t)firstChunk*32;i<n && i<(uint64_t)endChunk*32;i++)
        if (wedge[i]!=0xffffffffu)
          pendingList.push_back((uint32_t)i);
    const uint32_t nPend=(uint32_t)pendingList.size();
    const uint32_t b0=pass==1?firstChunk:0,b1=pass==1?endChunk:(nPend+31)/32;
    for (uint32_t b=b0;b<b1;b++)
    {
      auto body=[&]
      {
        if (pass==1)
          wb_classify_kernel<1>(sx,sy,sz,n,nChunks,bounds.data(),levelOff.data(),levelCnt.data(),nLevels,winner.data(),hyp,
                                maxSlope,thickness,clsIn.data(),perm.data(),0u,0xffffffffu,nullptr,labelSorted,counters,
                                wedge.data(),pending.data(),nullptr,0u);
        else
          wb_classify_kernel<2>(sx,sy,sz,n,nChunks,bounds.data(),levelOff.data(),levelCnt.data(),nLevels,winner.data(),hyp,
                                maxSlope,thickness,clsIn.data(),perm.data(),0u,0xffffffffu,nullptr,labelSorted,counters,
                                wedge.data(),pending.data(),pendingList.data(),nPend);
      };
      coll+=simt::run_warp(body,0,b,WB_CL_WARPS*32,nChunks);
    }
#else
    for (uint32_t b=firstChunk;b<endChunk;b++)
    {
      auto body=[&]
      {
        if (pass==1)
          wb_classify_kernel<1>(sx,sy,sz,n,nChunks,bounds.data(),levelOff.data(),levelCnt.data(),nLevels,winner.data(),hyp,
                                maxSlope,thickness,clsIn.data(),perm.data(),0u,0xffffffffu,nullptr,labelSorted,counters,
                                wedge.data(),pending.data());
        else
          wb_classify_kernel<2>(sx,sy,sz,n,nChunks,bounds.data(),levelOff.data(),levelCnt.data(),nLevels,winner.data(),hyp,
                                maxSlope,thickness,clsIn.data(),perm.data(),0u,0xffffffffu,nullptr,labelSorted,counters,
                                wedge.data(),pending.data());
      };
      coll+=simt::run_warp(body,0,b,WB_CL_WARPS*32,nChunks);
    }
#endif
  };
  static_assert(WB_CL_WARPS==1,"the emulator launches one warp per block");
  run(1);
  run(2);
  if (collectives)
    *collectives=coll;
  return 0;
}


// ------------------------------------------------------------------------------------------------------------
// The tile phases from the kernel sources: membership (wb_member_count_kernel, wb_member_fill_kernel), the sort by
// tile (a host stable sort stands in for the radix sort), wb_segment_kernel, wb_scan_kernel, then
// wb_tile_extent_list_kernel / wb_tile_grid_list_kernel / wb_postscan_list_kernel — launched in the order and with the grids of
// wb_scan and wb_postscan (wolken_b200.cu) — and finally the classify kernels with the membership's `winner` and
// the dense tile table, exactly the arrays the GPU path hands them.
struct SimtTile { int32_t n,nPoints,treeFlags,pad; double density,hyperboloidSize,height; };

extern "C" int simt_scan_classify(const double *sx,const double *sy,const double *sz,uint64_t n,
                                  const double cube[4],double tileSize,double minHyp,double maxSlope,double thickness,
                                  int doPostscan,
                                  SimtTile *tilesOut,uint64_t tilesCap,uint64_t *nTilesOut,   // non-empty tiles, ascending n
                                  uint8_t *labelSorted /* n bytes or NULL: skip classify */)
{
  {
    double t[512],co[512],si[512];
    wbhost::fillTanTables(t,co,si);
    memcpy(g_tanTable,t,sizeof(t)); memcpy(g_cosTable,co,sizeof(co)); memcpy(g_sinTable,si,sizeof(si));
    wbhost::fillFlowsnakeTables(g_fwdTable);
  }
  WbSnake snake;
  int lo,hi;
  double spacing;
  wbhost::snakeSetSize(cube[3],tileSize,&spacing,&lo,&hi);
  snake.spacing=spacing; snake.ccx=cube[0]; snake.ccy=cube[1]; snake.radius=spacing*41/71; snake.lo=lo; snake.hi=hi;
  const uint32_t T=(uint32_t)((long long)hi-lo+1);
  auto grid=[](uint64_t items,unsigned block) { return (unsigned)((items+block-1)/block); };
  // membership
  std::vector<uint32_t> cnt(n+1,0),off(n+1,0),winner(n);
  std::vector<uint4> tilesOf(n);
  launch(grid(n,128),128,[&]{ wb_member_count_kernel(sx,sy,n,snake,cnt.data(),tilesOf.data(),winner.data()); });
  for (uint64_t i=0;i<n;i++)
    off[i+1]=off[i]+cnt[i];
  const uint32_t m=off[n];
  std::vector<uint32_t> pairKey(m+1);
  std::vector<uint32_t> pairVal(m+1);
  launch(grid(n,256),256,[&]{ wb_member_fill_kernel(cnt.data(),off.data(),tilesOf.data(),n,pairKey.data(),pairVal.data()); });
  {
    std::vector<uint32_t> idx(m);
    for (uint32_t i=0;i<m;i++) idx[i]=i;
    std::stable_sort(idx.begin(),idx.end(),[&](uint32_t a,uint32_t b){ return pairKey[a]<pairKey[b]; });
    std::vector<uint32_t> k2(m+1);
    std::vector<uint32_t> v2(m+1);
    for (uint32_t i=0;i<m;i++) { k2[i]=pairKey[idx[i]]; v2[i]=pairVal[idx[i]]; }
    pairKey.swap(k2); pairVal.swap(v2);
  }
  std::vector<uint32_t> tStart(T,0),tCount(T,0),tileList(std::min<uint64_t>(T,m)+1);
  std::vector<int> tNPoints(T,0);
  std::vector<uint8_t> tTree(T,0);
  std::vector<double> tDensity(T,0),tHyp(T,0),tHeight(T,0);
  unsigned long long nList=0;
  if (m)
    launch(grid(m,256),256,[&]{ wb_segment_kernel(pairKey.data(),m,tStart.data(),tCount.data(),tileList.data(),&nList); });
  const int bundle=wb_scan_bundle(nList,m);
  if (nList)
    launch(grid(nList,WB_SCAN_WARPS*bundle),WB_SCAN_WARPS*32,[&]{ wb_scan_kernel(tileList.data(),(uint32_t)nList,tStart.data(),tCount.data(),
                                                          pairVal.data(),sx,sy,sz,snake,minHyp,bundle,tNPoints.data(),tTree.data(),
                                                          tDensity.data(),tHyp.data(),tHeight.data()); });
  if (doPostscan)
  {
    int ext[4]={INT_MAX,INT_MAX,INT_MIN,INT_MIN};
    const uint32_t nl=(uint32_t)nList;
    if (nl)
      launch(grid(nl,256),256,[&]{ wb_tile_extent_list_kernel(tileList.data(),nl,snake,-INFINITY,INFINITY,ext); });
    uint64_t cells=1;
    if (ext[0]<=ext[2])
      cells=(uint64_t)((long long)ext[2]-ext[0]+1)*(uint64_t)((long long)ext[3]-ext[1]+1);
    std::vector<uint8_t> tileGrid(cells,0);
    if (nl)
    {
      launch(grid(nl,256),256,[&]{ wb_tile_grid_list_kernel(tileList.data(),nl,tTree.data(),snake,-INFINITY,INFINITY,ext,tileGrid.data()); });
      launch(grid(nl,128),128,[&]{ wb_postscan_list_kernel(tileList.data(),nl,tTree.data(),snake,ext,tileGrid.data(),tHyp.data()); });
    }
  }
  uint64_t nt=0;
  for (uint32_t t=0;t<T;t++)
    if (tNPoints[t])
    {
      if (nt<tilesCap)
        tilesOut[nt]=SimtTile{(int32_t)((long long)t+lo),tNPoints[t],tTree[t],0,tDensity[t],tHyp[t],tHeight[t]};
      nt++;
    }
  *nTilesOut=nt;
  if (!labelSorted)
    return 0;
  // classify with the membership's winner and the dense table, as wb_classify does
  const uint32_t nChunks=(uint32_t)((n+31)/32);
  std::vector<uint32_t> levelOff,levelCnt;
  uint64_t total=0;
  uint32_t c=nChunks;
  while (true)
  {
    levelOff.push_back((uint32_t)total); levelCnt.push_back(c); total+=c;
    if (c<=32) break;
    c=(c+31)/32;
  }
  const int nLevels=(int)levelCnt.size();
  std::vector<WbBound> bounds(total);
  levelOff.resize(16); levelCnt.resize(16);
  launch(grid((uint64_t)nChunks*32,256),256,[&]{ wb_chunk_bounds_kernel(sx,sy,sz,n,bounds.data(),nChunks); });
  for (int l=1;l<nLevels;l++)
    launch(grid((uint64_t)levelCnt[l]*32,256),256,[&]{ wb_node_bounds_kernel(bounds.data()+levelOff[l-1],levelCnt[l-1],
                                                                           bounds.data()+levelOff[l],levelCnt[l]); });
  std::vector<uint32_t> perm(n),wedge(n,0xffffffffu);
  std::vector<uint8_t> clsIn(n,0),pending(nChunks,0);
  for (uint64_t i=0;i<n;i++) perm[i]=(uint32_t)i;
  unsigned long long counters[24]={0};
  for (int pass=1;pass<=2;pass++)
  {
    uint32_t nBlocks=nChunks;
#if WB_CL_COMPACT2
    // pass 2 walks the pending queries gathered into full warps, as wb_classify does (flag, scan, scatter)
    std::vector<uint32_t> pendingList;
    if (pass==2)
    {
      for (uint64_t i=0;i<n;i++)
        if (wedge[i]!=0xffffffffu)
          pendingList.push_back((uint32_t)i);
      nBlocks=(uint32_t)((pendingList.size()+31)/32);
    }
    const uint32_t nPend=(uint32_t)pendingList.size();
#endif
    for (uint32_t b=0;b<nBlocks;b++)
      simt::run_warp([&]
      {
        if (pass==1)
          wb_classify_kernel<1>(sx,sy,sz,n,nChunks,bounds.data(),levelOff.data(),levelCnt.data(),nLevels,winner.data(),tHyp.data(),
                                maxSlope,thickness,clsIn.data(),perm.data(),0u,0xffffffffu,nullptr,labelSorted,counters,wedge.data(),pending.data()
#if WB_CL_COMPACT2
                                ,nullptr,0u
#endif
                                );
        else
          wb_classify_kernel<2>(sx,sy,sz,n,nChunks,bounds.data(),levelOff.data(),levelCnt.data(),nLevels,winner.data(),tHyp.data(),
                                maxSlope,thickness,clsIn.data(),perm.data(),0u,0xffffffffu,nullptr,labelSorted,counters,wedge.data(),pending.data()
#if WB_CL_COMPACT2
                                ,pendingList.data(),nPend
#endif
                                );
      },0,b,WB_CL_WARPS*32,nBlocks);
  }
  return 0;
}


// ------------------------------------------------------------------------------------------------------------
// The radix sort (wb_sort.cuh) with whole blocks emulated (its kernels meet at __syncthreads): the launch sequence of
// wb_exclusive_scan and wb_radix_sort, replayed.
template <typename F> static void launchBlocks(unsigned nBlocks,unsigned blockThreads,F body)
{
  for (unsigned b=0;b<nBlocks;b++)
    simt::run_block(body,blockThreads,0,b,blockThreads,nBlocks);
}

static void emuExclusiveScan(const uint32_t *in,uint32_t *out,uint64_t n,std::vector<uint32_t> &blockSums)
{
  if (!n)
    return;
  const uint64_t nb=(n+WB_SCAN_TILE-1)/WB_SCAN_TILE;
  blockSums.assign(nb+1024,0);
  launchBlocks((unsigned)nb,WB_SCAN_THREADS,[&]{ wb_scan_reduce_kernel(in,n,blockSums.data()); });
  launchBlocks(1,1024,[&]{ wb_scan_single_kernel(blockSums.data(),(uint32_t)nb,nullptr); });
  launchBlocks((unsigned)nb,WB_SCAN_THREADS,[&]{ wb_scan_apply_kernel(in,out,n,blockSums.data()); });
}

extern "C" int simt_radix_sort(uint64_t *keys,uint32_t *vals,uint64_t n,int beginBit,int endBit)
// stable sort of (key,val) by key bits [beginBit,endBit), in place (the result is copied back if it ends in the
// second buffer)
{
  if (!n)
    return 0;
  std::vector<uint64_t> kb(n);
  std::vector<uint32_t> vb(n),table,blockSums;
  const uint64_t nb=(n+WB_SORT_TILE-1)/WB_SORT_TILE;
  table.assign(nb*256+256,0);
  uint64_t *ki=keys,*ko=kb.data();
  uint32_t *vi=vals,*vo=vb.data();
  for (int shift=beginBit;shift<endBit;shift+=8)
  {
    launchBlocks((unsigned)nb,WB_SORT_THREADS,[&]{ wb_sort_upsweep_kernel(ki,n,shift,table.data(),(uint32_t)nb); });
    emuExclusiveScan(table.data(),table.data(),nb*256,blockSums);
    launchBlocks((unsigned)nb,WB_SORT_THREADS,[&]{ wb_sort_downsweep_kernel(ki,vi,ko,vo,n,shift,table.data(),(uint32_t)nb); });
    std::swap(ki,ko);
    std::swap(vi,vo);
  }
  if (ki!=keys)
  {
    memcpy(keys,ki,n*sizeof(uint64_t));
    memcpy(vals,vi,n*sizeof(uint32_t));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// LAS decode (wb_decode_kernel: a CTA stages its byte span in shared memory and meets at __syncthreads).
extern "C" int simt_decode(const uint8_t *recs,uint64_t n,int fmt,int recLen,int dropZeros,int misalign,
                           int32_t *xi,int32_t *yi,int32_t *zi,uint8_t *cls,uint8_t *ret,unsigned long long *nDropped)
// misalign (0..15): where the first record starts relative to a 16-byte boundary (the kernel loads aligned vectors
// that may begin before its span)
{
  std::vector<uint8_t> buf(n*recLen+64);
  uint8_t *base=(uint8_t *)(((uintptr_t)buf.data()+31)&~(uintptr_t)15)+(misalign&15);
  memcpy(base,recs,n*recLen);
  *nDropped=0;
  launchBlocks((unsigned)((n+WB_DEC_THREADS-1)/WB_DEC_THREADS),WB_DEC_THREADS,
               [&]{ wb_decode_kernel(base,n,fmt,recLen,dropZeros,xi,yi,zi,cls,ret,nDropped); });
  return 0;
}
