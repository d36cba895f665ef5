// wolken_b200.cu — C ABI of libwolken_b200.so (see include/wolken_b200.h): context, device
// memory, phase drivers and timing around the kernels in wb_kernels.cuh.
#include <cstdio>
#include <climits>
#include <cstdarg>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include "../../include/wolken_b200.h"
#include <thread>
#include <mutex>
#include <condition_variable>
#include <fcntl.h>
#include <unistd.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include "wb_host.h"
#include "wb_kernels.cuh"
#include "wb_sort.cuh"
#include "wb_encode.cuh"
#include "wb_query.cuh"

namespace
{

#define WB_READ_THREADS 8          // pread workers = pinned buffers of wb_add_las_file (4: 17.7 GB/s from the page cache)

template <typename T> struct DevBuf
{
  T *p=nullptr;
  uint64_t cap=0;
  cudaError_t ensure(uint64_t n)
  {
    if (n<=cap)
      return cudaSuccess;
    if (p)
      cudaFree(p);
    p=nullptr;
    cap=0;
    cudaError_t e=cudaMalloc((void **)&p,(size_t)(n*sizeof(T)));
    if (e==cudaSuccess)
      cap=n;
    return e;
  }
  void release()
  {
    if (p)
      cudaFree(p);
    p=nullptr;
    cap=0;
  }
};
template <typename T> struct TmpBuf:DevBuf<T>      // a scratch buffer of one call
{
  ~TmpBuf() { this->release(); }
};

enum Phase { PH_EMPTY=0,PH_LOADED,PH_BUILT,PH_SCANNED,PH_POSTSCANNED,PH_CLASSIFIED };

} // namespace

struct WbShard;                                  // wb_shard.cuh: state of the sharded pipeline
static void wbShardFree(wb_ctx *ctx);

struct wb_ctx
{
  int device=0;
  std::string err;
  cudaStream_t st=nullptr,stCopy=nullptr,stLoad=nullptr;   // stLoad: decode of arriving records, highest priority
  cudaEvent_t evA=nullptr,evB=nullptr,evC=nullptr,evD=nullptr,evCopy[2]={nullptr,nullptr},evDec[2]={nullptr,nullptr};
  cudaEvent_t evMark[WB_MARKS]={},evJoin=nullptr;     // wb_mark: caller-placed timing events on the compute stream
  WbParams prm{1,1,0,0.1};
  Phase phase=PH_EMPTY;
  // cloud
  uint64_t n=0,nValid=0,reserved=0;
  std::vector<WbSegment> segs;
  std::vector<double> corners;
  bool geomOverride=false;
  wb_geometry geom{};
  WbSnake snake{};
  // device arrays (input order)
  DevBuf<int> xi,yi,zi;
  DevBuf<uint8_t> cls,ret,labelIn,labelSorted,leafDepth;
  DevBuf<unsigned long long> keyA,keyB,counters;
  DevBuf<uint32_t> pairKeyA,pairKeyB;     // tile number of every (point, covering tile) pair: 32 bits, a third less sort traffic
  DevBuf<uint32_t> idxA,idxB,scr0,scr1,winner,pairValA,pairValB,table,blockSums;
  DevBuf<uint32_t> dupIn,dupRep;          // input indices of (lost duplicate, surviving point at the same XYZ)
  uint64_t nDup=0;
  DevBuf<uint8_t> forcedIn;               // sharded runs: halo points that stand in for an own record at the same XYZ
  bool haveForced=false;
  WbShard *shard=nullptr;
  // pinned ring of the file reader (wb_add_las_file)
  uint8_t *readBuf[WB_READ_THREADS]={};
  uint64_t readBufBytes=0;
  cudaEvent_t evRead[WB_READ_THREADS]={};
  // raw records kept for wb_encode (one buffer per wb_add_las call)
  bool keepRecords=false;
  bool keepZeroReturns=false;          // wb_set_return_zero_rule(ctx,1)
  std::vector<uint8_t *> recBufs;
  std::vector<WbRecSeg> recSegs;
  DevBuf<WbRecSegs> drsegs;
  DevBuf<uint8_t> outArena,encLut;
  uint64_t outBytes=0;                     // valid bytes in outArena after wb_encode
  DevBuf<unsigned long long> encDest,encCount;
  DevBuf<uint32_t> encFile,encCounts,attrSrc,invPerm;
  DevBuf<int> encMinMax;
  DevBuf<uint4> tilesOf;
  DevBuf<double> sx,sy,sz;
  DevBuf<int4> packed;                    // (X, Y, Z, index of the file's header) per input point, for the gather
  DevBuf<uint8_t> staging[2];
  DevBuf<WbSegments> dsegs;
  DevBuf<WbNode> nodesA,nodesB;
  DevBuf<WbLeafDev> leaves;
  DevBuf<WbBound> bounds;
  DevBuf<uint32_t> levelOff,levelCnt;
  // tiles
  uint32_t nTiles=0;
  DevBuf<uint32_t> tStart,tCount,tileList;
  DevBuf<int> tNPoints;
  DevBuf<uint8_t> tTree;
  DevBuf<double> tDensity,tHyp,tHeight;
  DevBuf<int> tileExt;
  DevBuf<uint32_t> wedgeBuf;
#if WB_CL_COMPACT2
  DevBuf<uint32_t> pendingList;
#endif
  DevBuf<uint8_t> chunkPending;
  DevBuf<uint8_t> tileGrid;
  // classify order (Hilbert over xy): sort scratch, the store's columns in that order, the hierarchy over them
  DevBuf<unsigned long long> hKeyA,hKeyB;
  DevBuf<uint32_t> hIdxA,hIdxB,hWinner,hPerm;
  DevBuf<double> hx,hy,hz;
  DevBuf<uint8_t> hLabel;
  DevBuf<WbBound> hBounds;
  // results of build
  unsigned long long *keys=nullptr;   // sorted keys (keyA or keyB)
  uint32_t *perm=nullptr;             // sorted -> input
  uint32_t *pairKeys=nullptr;
  uint32_t *pairVals=nullptr;
  uint64_t nPairs=0;
  uint32_t nLeaves=0,nChunks=0;
  uint32_t ownFirst=0,ownEnd=0xffffffffu;   // input-index range this GPU labels (the rest is halo)
  int nLevels=0;
  std::vector<uint32_t> hLevelOff,hLevelCnt;
  wb_stats stats{};
  bool tablesUploaded=false;
  bool pacedCopies=false;             // WB_H2D_PACED=1: wb_add_las keeps at most two chunk copies queued (no gain measured)
  bool windowOn=false;                // wb_set_window: keep only the records with x in [windowLo,windowHi)
  double windowLo=0,windowHi=0;
  bool storeHilbert=false;            // the store is a classify-only one (buildStore(ctx,true)): Hilbert order, no leaves
};

static int ensureTileArrays(wb_ctx *ctx);

namespace
{

int fail(wb_ctx *c,int code,const char *fmt,...)
{
  char buf[512];
  va_list ap;
  va_start(ap,fmt);
  vsnprintf(buf,sizeof(buf),fmt,ap);
  va_end(ap);
  if (c)
    c->err=buf;
  return code;
}

#define CK(call) do { cudaError_t e_=(call); if (e_!=cudaSuccess) return fail(ctx,WB_ERR_CUDA,"%s: %s (%s:%d)",#call,cudaGetErrorString(e_),__FILE__,__LINE__); } while (0)
#define KCHECK() CK(cudaGetLastError())

inline unsigned gridFor(uint64_t n,unsigned block) { return (unsigned)((n+block-1)/block); }

float elapsed(cudaEvent_t a,cudaEvent_t b)
{
  float ms=0;
  cudaEventElapsedTime(&ms,a,b);
  return ms;
}

int uploadTables(wb_ctx *ctx)
// The tables are __device__ globals: once per DEVICE, not per context (a second context of the same device must not
// rewrite them under the kernels of the first).
{
  static std::mutex m;
  static bool done[256]={};
  if (ctx->tablesUploaded)
    return WB_OK;
  std::lock_guard<std::mutex> lk(m);
  if (ctx->device>=0 && ctx->device<256 && done[ctx->device])
  {
    ctx->tablesUploaded=true;
    return WB_OK;
  }
  double t[512],co[512],si[512];
  unsigned char fw[96];
  wbhost::fillTanTables(t,co,si);
  wbhost::fillFlowsnakeTables(fw);
  CK(cudaMemcpyToSymbol(g_tanTable,t,sizeof(t)));
  CK(cudaMemcpyToSymbol(g_cosTable,co,sizeof(co)));
  CK(cudaMemcpyToSymbol(g_sinTable,si,sizeof(si)));
  CK(cudaMemcpyToSymbol(g_fwdTable,fw,sizeof(fw)));
  if (ctx->device>=0 && ctx->device<256)
    done[ctx->device]=true;
  ctx->tablesUploaded=true;
  return WB_OK;
}

template <typename T> int growKeep(wb_ctx *ctx,DevBuf<T> &b,uint64_t keep,uint64_t n)
// enlarge b to n elements, preserving the first `keep`
{
  if (n<=b.cap)
    return WB_OK;
  T *np=nullptr;
  CK(cudaMalloc((void **)&np,(size_t)(n*sizeof(T))));
  if (keep && b.p)
    CK(cudaMemcpy(np,b.p,(size_t)(keep*sizeof(T)),cudaMemcpyDeviceToDevice));
  if (b.p)
    cudaFree(b.p);
  b.p=np;
  b.cap=n;
  return WB_OK;
}

int ensurePointArrays(wb_ctx *ctx,uint64_t n)
{
  if (n<=ctx->xi.cap)
    return WB_OK;
  uint64_t want=ctx->n?std::max<uint64_t>(n,ctx->xi.cap*2):n;      // grow geometrically once files are in
  int rc;
  CK(cudaStreamSynchronize(ctx->st));
  if ((rc=growKeep(ctx,ctx->xi,ctx->n,want)) || (rc=growKeep(ctx,ctx->yi,ctx->n,want)) ||
      (rc=growKeep(ctx,ctx->zi,ctx->n,want)) || (rc=growKeep(ctx,ctx->cls,ctx->n,want)) ||
      (rc=growKeep(ctx,ctx->ret,ctx->n,want)))
    return rc;
  ctx->reserved=want;
  return WB_OK;
}

int computeGeometry(wb_ctx *ctx)
{
  wb_geometry &g=ctx->geom;
  if (!ctx->geomOverride)
  {
    if (ctx->corners.empty())
      return fail(ctx,WB_ERR_STATE,"no extents: call wb_add_extent (or wb_set_geometry) first");
    wbhost::sizeFit(ctx->corners.data(),(int)(ctx->corners.size()/3),g.root_center,&g.root_side);
    wbhost::bboxCube(ctx->corners.data(),(int)(ctx->corners.size()/3),g.cube);
  }
  if (!(g.root_side>0) || !(g.cube[3]>0))
    return fail(ctx,WB_ERR_ARG,"degenerate extents (side %g, cube %g)",g.root_side,g.cube[3]);
  int lo,hi;
  g.snake_index=wbhost::snakeSetSize(g.cube[3],ctx->prm.tileSize,&g.spacing,&lo,&hi);
  g.snake_lo=lo;
  g.snake_hi=hi;
  g.radius=g.spacing*41/71;                        // Flowsnake::cyl, flowsnake.cpp:265
  ctx->snake.spacing=g.spacing;
  ctx->snake.ccx=g.cube[0];
  ctx->snake.ccy=g.cube[1];
  ctx->snake.radius=g.radius;
  ctx->snake.lo=lo;
  ctx->snake.hi=hi;
  ctx->nTiles=(uint32_t)((long long)hi-lo+1);
  return WB_OK;
}

int decodeDevice(wb_ctx *ctx,const uint8_t *d,uint64_t first,uint64_t cnt,int fmt,int recLen,int dropZeros,cudaStream_t st)
{
  if (!cnt)
    return WB_OK;
  wb_decode_kernel<<<gridFor(cnt,WB_DEC_THREADS),WB_DEC_THREADS,0,st>>>(
      d,cnt,fmt,recLen,dropZeros,ctx->xi.p+first,ctx->yi.p+first,ctx->zi.p+first,
      ctx->cls.p+first,ctx->ret.p+first,ctx->counters.p+2);
  ctx->stats.kernel_launches++;
  KCHECK();
  return WB_OK;
}

int applyWindow(wb_ctx *ctx,uint64_t n,const double scale[3],const double offset[3],double unit,uint64_t *kept)
// The n records just decoded at [ctx->n,ctx->n+n): keep those inside the context's x-window, in their order.
{
  *kept=n;
  if (!ctx->windowOn || !n)
    return WB_OK;
  if (ctx->keepRecords)
    return fail(ctx,WB_ERR_STATE,"wb_set_window and wb_keep_records do not go together");
  cudaStream_t st=ctx->st;
  const uint64_t first=ctx->n;
  WbSegment sg;
  sg.first=first;
  sg.count=n;
  for (int k=0;k<3;k++)
  {
    sg.scale[k]=scale[k];
    sg.offset[k]=offset[k];
  }
  sg.unit=unit;
  TmpBuf<uint32_t> flag,off;
  TmpBuf<int> tx,ty,tz;
  TmpBuf<uint8_t> tc,tr;
  CK(flag.ensure(n+1)); CK(off.ensure(n+1));
  CK(ctx->blockSums.ensure(wb_div_up(n+1,WB_SCAN_TILE)+1024));
  CK(cudaMemsetAsync(flag.p+n,0,sizeof(uint32_t),st));
  wb_window_flag_kernel<<<gridFor(n,256),256,0,st>>>(ctx->xi.p,ctx->ret.p,first,n,sg,ctx->windowLo,ctx->windowHi,flag.p,
                                                    ctx->counters.p+2);
  ctx->stats.kernel_launches++;
  KCHECK();
  CK(wb_exclusive_scan(flag.p,off.p,n+1,ctx->blockSums.p,ctx->blockSums.cap,st,&ctx->stats.kernel_launches));
  uint32_t m=0;
  CK(cudaMemcpyAsync(&m,off.p+n,sizeof(uint32_t),cudaMemcpyDeviceToHost,st));
  CK(cudaStreamSynchronize(st));
  if (m<n)
  {
    CK(tx.ensure(m+1)); CK(ty.ensure(m+1)); CK(tz.ensure(m+1)); CK(tc.ensure(m+1)); CK(tr.ensure(m+1));
    wb_window_scatter_kernel<<<gridFor(n,256),256,0,st>>>(flag.p,off.p,first,n,ctx->xi.p,ctx->yi.p,ctx->zi.p,ctx->cls.p,ctx->ret.p,
                                                         tx.p,ty.p,tz.p,tc.p,tr.p);
    if (m)
      wb_window_copy_back_kernel<<<gridFor(m,256),256,0,st>>>(tx.p,ty.p,tz.p,tc.p,tr.p,m,first,ctx->xi.p,ctx->yi.p,ctx->zi.p,
                                                             ctx->cls.p,ctx->ret.p);
    ctx->stats.kernel_launches+=2;
    KCHECK();
    CK(cudaStreamSynchronize(st));
  }
  *kept=m;
  return WB_OK;
}

int addSegment(wb_ctx *ctx,uint64_t n,const double scale[3],const double offset[3],double unit,
               const uint8_t *keptRecs=nullptr,int fmt=-1,int recLen=0)
{
  if (ctx->segs.size()>=WB_MAX_SEGMENTS)
    return fail(ctx,WB_ERR_ARG,"too many input files (max %d)",WB_MAX_SEGMENTS);
  WbRecSeg rs;
  rs.recs=keptRecs;
  rs.fmt=fmt;
  rs.recLen=recLen;
  ctx->recSegs.push_back(rs);
  WbSegment s;
  s.first=ctx->n;
  s.count=n;
  for (int k=0;k<3;k++)
  {
    s.scale[k]=scale[k];
    s.offset[k]=offset[k];
  }
  s.unit=unit;
  ctx->segs.push_back(s);
  ctx->n+=n;
  ctx->phase=PH_LOADED;
  return WB_OK;
}

void freeKeptRecords(wb_ctx *ctx)
{
  for (uint8_t *b:ctx->recBufs)
    cudaFree(b);
  ctx->recBufs.clear();
  ctx->recSegs.clear();
}

int checkFormat(wb_ctx *ctx,int fmt,int recLen)
{
  static const int len[11]={20,28,26,34,57,63,30,36,38,59,67};   // las.cpp:38
  if (fmt<0 || fmt>10)
    return fail(ctx,WB_ERR_FORMAT,"point format %d unknown",fmt);
  if (recLen<len[fmt])
    return fail(ctx,WB_ERR_FORMAT,"record length %d shorter than format %d needs (%d)",recLen,fmt,len[fmt]);
  if (recLen>WB_DEC_MAXLEN)
    return fail(ctx,WB_ERR_FORMAT,"record length %d not supported (waveform formats 4,5,9,10 are out of scope)",recLen);
  return WB_OK;
}

} // namespace

// ============================================================================ life cycle

extern "C" int wb_device_count(int *out)
{
  if (!out)
    return WB_ERR_ARG;
  int ndev=0;
  if (cudaGetDeviceCount(&ndev)!=cudaSuccess)
    ndev=0;
  *out=ndev;
  return ndev>0?WB_OK:WB_ERR_CUDA;
}

extern "C" int wb_create(int device,wb_ctx **out)
{
  if (!out)
    return WB_ERR_ARG;
  *out=nullptr;
  int ndev=0;
  if (cudaGetDeviceCount(&ndev)!=cudaSuccess || ndev==0 || device<0 || device>=ndev)
    return WB_ERR_CUDA;               // no CPU fallback
  wb_ctx *ctx=new wb_ctx;
  ctx->device=device;
  if (cudaSetDevice(device)!=cudaSuccess)
  {
    delete ctx;
    return WB_ERR_CUDA;
  }
  cudaStreamCreateWithFlags(&ctx->st,cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&ctx->stCopy,cudaStreamNonBlocking);
  {
    // Records may arrive for one context while ANOTHER context of the same device is deep in classify (a job that
    // streams clouds through two contexts, bench.py's e2e): the decode kernels are tiny but would queue behind the
    // millions of blocks classify still has pending, and the H2D pipeline with them.  Highest priority lets their
    // blocks take the next free slots.
    int least=0,greatest=0;
    cudaDeviceGetStreamPriorityRange(&least,&greatest);
    cudaStreamCreateWithPriority(&ctx->stLoad,cudaStreamNonBlocking,greatest);
    cudaEventCreateWithFlags(&ctx->evJoin,cudaEventDisableTiming);
    const char *paced=getenv("WB_H2D_PACED");
    ctx->pacedCopies=paced && paced[0]=='1';
  }
  cudaEventCreate(&ctx->evA); cudaEventCreate(&ctx->evB); cudaEventCreate(&ctx->evC); cudaEventCreate(&ctx->evD);
  for (int i=0;i<2;i++)
  {
    cudaEventCreateWithFlags(&ctx->evCopy[i],cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->evDec[i],cudaEventDisableTiming);
  }
  if (ctx->counters.ensure(24)!=cudaSuccess || ctx->dsegs.ensure(1)!=cudaSuccess)
  {
    delete ctx;
    return WB_ERR_CUDA;
  }
  cudaMemset(ctx->counters.p,0,24*sizeof(unsigned long long));
  if (uploadTables(ctx)!=WB_OK)
  {
    delete ctx;
    return WB_ERR_CUDA;
  }
  *out=ctx;
  return WB_OK;
}

extern "C" void wb_destroy(wb_ctx *ctx)
{
  if (!ctx)
    return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  ctx->xi.release(); ctx->yi.release(); ctx->zi.release(); ctx->cls.release(); ctx->ret.release();
  ctx->labelIn.release(); ctx->labelSorted.release(); ctx->leafDepth.release();
  ctx->keyA.release(); ctx->keyB.release(); ctx->pairKeyA.release(); ctx->pairKeyB.release(); ctx->counters.release();
  ctx->idxA.release(); ctx->idxB.release(); ctx->scr0.release(); ctx->scr1.release(); ctx->winner.release();
  ctx->pairValA.release(); ctx->pairValB.release(); ctx->table.release(); ctx->blockSums.release();
  ctx->tilesOf.release(); ctx->sx.release(); ctx->sy.release(); ctx->sz.release();
  ctx->staging[0].release(); ctx->staging[1].release(); ctx->dsegs.release();
  ctx->nodesA.release(); ctx->nodesB.release(); ctx->leaves.release(); ctx->bounds.release();
  ctx->levelOff.release(); ctx->levelCnt.release(); ctx->packed.release();
  ctx->hKeyA.release(); ctx->hKeyB.release(); ctx->hIdxA.release(); ctx->hIdxB.release(); ctx->hWinner.release(); ctx->hPerm.release();
  ctx->hx.release(); ctx->hy.release(); ctx->hz.release(); ctx->hLabel.release(); ctx->hBounds.release();
  ctx->tStart.release(); ctx->tCount.release(); ctx->tileList.release(); ctx->tNPoints.release(); ctx->tTree.release();
  ctx->tDensity.release(); ctx->tHyp.release(); ctx->tHeight.release(); ctx->tileExt.release(); ctx->tileGrid.release(); ctx->wedgeBuf.release(); ctx->chunkPending.release();
#if WB_CL_COMPACT2
  ctx->pendingList.release();
#endif
 
  ctx->dupIn.release(); ctx->dupRep.release(); ctx->forcedIn.release();
  wbShardFree(ctx);
  freeKeptRecords(ctx);
  for (int i=0;i<WB_READ_THREADS;i++)
  {
    if (ctx->readBuf[i]) cudaFreeHost(ctx->readBuf[i]);
    if (ctx->evRead[i]) cudaEventDestroy(ctx->evRead[i]);
  }
  ctx->drsegs.release(); ctx->outArena.release(); ctx->encLut.release(); ctx->encDest.release(); ctx->encCount.release();
  ctx->encFile.release(); ctx->encCounts.release(); ctx->attrSrc.release(); ctx->invPerm.release(); ctx->encMinMax.release();
  cudaStreamDestroy(ctx->st); cudaStreamDestroy(ctx->stCopy); cudaStreamDestroy(ctx->stLoad);
  cudaEventDestroy(ctx->evA); cudaEventDestroy(ctx->evB); cudaEventDestroy(ctx->evC); cudaEventDestroy(ctx->evD);
  for (int i=0;i<2;i++)
  {
    cudaEventDestroy(ctx->evCopy[i]);
    cudaEventDestroy(ctx->evDec[i]);
  }
  for (int i=0;i<WB_MARKS;i++)
    if (ctx->evMark[i]) cudaEventDestroy(ctx->evMark[i]);
  if (ctx->evJoin) cudaEventDestroy(ctx->evJoin);
  delete ctx;
}

extern "C" const char *wb_last_error(wb_ctx *ctx)
{
  return ctx?ctx->err.c_str():"null context";
}

extern "C" int wb_reserve(wb_ctx *ctx,uint64_t n)
{
  if (!ctx)
    return WB_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (ctx->n)
    return fail(ctx,WB_ERR_STATE,"wb_reserve must precede wb_add_las");
  return ensurePointArrays(ctx,n);
}

extern "C" int wb_clear(wb_ctx *ctx)
{
  if (!ctx)
    return WB_ERR_ARG;
  cudaSetDevice(ctx->device);
  CK(cudaStreamSynchronize(ctx->st));
  ctx->n=ctx->nValid=0;
  ctx->nDup=0;
  freeKeptRecords(ctx);
  ctx->outArena.release();            // sized by the last output; may be gigabytes
  ctx->outBytes=0;
  ctx->segs.clear();
  ctx->corners.clear();
  ctx->geomOverride=false;
  ctx->phase=PH_EMPTY;
  ctx->nPairs=0;
  ctx->nLeaves=0;
  ctx->ownFirst=0;
  ctx->ownEnd=0xffffffffu;
  ctx->windowOn=false;
  uint64_t launches=ctx->stats.kernel_launches;
  memset(&ctx->stats,0,sizeof(ctx->stats));
  ctx->stats.kernel_launches=launches;
  CK(cudaMemsetAsync(ctx->counters.p,0,24*sizeof(unsigned long long),ctx->st));
  return WB_OK;
}

extern "C" int wb_set_params(wb_ctx *ctx,double tileSize,double maxSlope,double thickness,double minHyp)
{
  if (!ctx)
    return WB_ERR_ARG;
  if (!(tileSize>0) || !(maxSlope>0) || !(minHyp>=0) || !std::isfinite(thickness))
    return fail(ctx,WB_ERR_ARG,"parameters out of range (tileSize>0, maxSlope>0, minHyperboloidSize>=0)");
  ctx->prm.tileSize=tileSize;
  ctx->prm.maxSlope=maxSlope;
  ctx->prm.thickness=thickness;
  ctx->prm.minHyp=minHyp;
  return WB_OK;
}

// ============================================================================ read

extern "C" int wb_add_extent(wb_ctx *ctx,const double mn[3],const double mx[3])
{
  if (!ctx || !mn || !mx)
    return WB_ERR_ARG;
  for (int k=0;k<3;k++)
    ctx->corners.push_back(mn[k]);
  for (int k=0;k<3;k++)
    ctx->corners.push_back(mx[k]);
  return WB_OK;
}

extern "C" int wb_set_geometry(wb_ctx *ctx,const double rc[3],double rs,const double cube[4])
{
  if (!ctx || !rc || !cube)
    return WB_ERR_ARG;
  for (int k=0;k<3;k++)
    ctx->geom.root_center[k]=rc[k];
  ctx->geom.root_side=rs;
  for (int k=0;k<4;k++)
    ctx->geom.cube[k]=cube[k];
  ctx->geomOverride=true;
  return WB_OK;
}

extern "C" int wb_get_geometry(wb_ctx *ctx,wb_geometry *out)
{
  if (!ctx || !out)
    return WB_ERR_ARG;
  int rc=computeGeometry(ctx);
  if (rc)
    return rc;
  *out=ctx->geom;
  return WB_OK;
}

extern "C" int wb_add_las_device(wb_ctx *ctx,const uint8_t *d,uint64_t n,int fmt,int recLen,
                                 const double scale[3],const double offset[3],double unit)
{
  if (!ctx || (!d && n) || !scale || !offset)
    return WB_ERR_ARG;
  cudaSetDevice(ctx->device);
  int rc=checkFormat(ctx,fmt,recLen);
  if (rc)
    return rc;
  if (ctx->phase>PH_LOADED)
    return fail(ctx,WB_ERR_STATE,"cloud already built: wb_clear first");
  if ((rc=ensurePointArrays(ctx,ctx->n+n)))
    return rc;
  int dropZeros=0;
  if (n)
  {
    uint8_t b14;
    CK(cudaMemcpyAsync(&b14,d+14,1,cudaMemcpyDeviceToHost,ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    dropZeros=(fmt<6?(b14&7):(b14&15))!=0;          // threads.cpp:485-500: decided by record 0
  }
  if (ctx->keepZeroReturns)
    dropZeros=0;
  CK(cudaEventRecord(ctx->evA,ctx->st));
  if ((rc=decodeDevice(ctx,d,ctx->n,n,fmt,recLen,dropZeros,ctx->st)))
    return rc;
  CK(cudaEventRecord(ctx->evB,ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  ctx->stats.ms_decode+=elapsed(ctx->evA,ctx->evB);
  uint8_t *kept=nullptr;
  if (ctx->keepRecords && n)
  {
    CK(cudaMalloc((void **)&kept,(size_t)(n*recLen+64)));
    ctx->recBufs.push_back(kept);
    CK(cudaMemcpyAsync(kept,d,(size_t)(n*recLen),cudaMemcpyDeviceToDevice,ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
  }
  {
    uint64_t inside=n;
    if ((rc=applyWindow(ctx,n,scale,offset,unit,&inside)))
      return rc;
    return addSegment(ctx,inside,scale,offset,unit,kept,fmt,recLen);
  }
}

extern "C" int wb_add_las(wb_ctx *ctx,const uint8_t *recs,uint64_t n,int fmt,int recLen,
                          const double scale[3],const double offset[3],double unit)
{
  if (!ctx || (!recs && n) || !scale || !offset)
    return WB_ERR_ARG;
  cudaSetDevice(ctx->device);
  int rc=checkFormat(ctx,fmt,recLen);
  if (rc)
    return rc;
  if (ctx->phase>PH_LOADED)
    return fail(ctx,WB_ERR_STATE,"cloud already built: wb_clear first");
  if ((rc=ensurePointArrays(ctx,ctx->n+n)))
    return rc;
  int dropZeros=n?((fmt<6?(recs[14]&7):(recs[14]&15))!=0):0;
  if (ctx->keepZeroReturns)
    dropZeros=0;
  // double-buffered pipeline: copy chunk i+1 on the copy stream while chunk i is decoded
  const uint64_t chunkRecs=(uint64_t)1<<21;          // 2 Mi records (multiple of 16: chunks stay 16-byte aligned)
  const uint64_t chunkBytes=chunkRecs*recLen;
  uint8_t *kept=nullptr;
  if (ctx->keepRecords && n)
  { // the records stay resident for wb_encode: copy straight to their final place and decode there
    CK(cudaMalloc((void **)&kept,(size_t)(n*recLen+64)));
    ctx->recBufs.push_back(kept);
  }
  else
    for (int b=0;b<2;b++)
      CK(ctx->staging[b].ensure(std::min<uint64_t>(chunkBytes,n*recLen)+64));
  cudaStream_t ld=ctx->stLoad;                       // after whatever the context's own stream still has queued
  CK(cudaEventRecord(ctx->evJoin,ctx->st));
  CK(cudaStreamWaitEvent(ld,ctx->evJoin,0));
  CK(cudaEventRecord(ctx->evA,ld));
  uint64_t done=0;
  int b=0,used[2]={0,0};
  while (done<n)
  {
    uint64_t cnt=std::min(chunkRecs,n-done);
    uint8_t *dst=kept?kept+done*recLen:ctx->staging[b].p;
    // Optional pacing (at most two chunk copies queued).  The copy engine takes its work in order, so a context that
    // classifies on this device meanwhile must keep its own host->device traffic to a few KB per call (inlined by the
    // driver): see the segment-table upload in buildStore.  With that in place pacing changed nothing at 8 GPUs.
    if (used[b] && ctx->pacedCopies)
      CK(cudaEventSynchronize(ctx->evCopy[b]));
    if (used[b] && !kept)
      CK(cudaStreamWaitEvent(ctx->stCopy,ctx->evDec[b],0));   // staging[b] free again?
    CK(cudaMemcpyAsync(dst,recs+done*recLen,cnt*recLen,cudaMemcpyHostToDevice,ctx->stCopy));
    CK(cudaEventRecord(ctx->evCopy[b],ctx->stCopy));
    CK(cudaStreamWaitEvent(ld,ctx->evCopy[b],0));
    if ((rc=decodeDevice(ctx,dst,ctx->n+done,cnt,fmt,recLen,dropZeros,ld)))
      return rc;
    CK(cudaEventRecord(ctx->evDec[b],ld));
    used[b]=1;
    done+=cnt;
    b^=1;
  }
  CK(cudaEventRecord(ctx->evB,ld));
  CK(cudaStreamSynchronize(ld));
  ctx->stats.ms_h2d+=elapsed(ctx->evA,ctx->evB);       // copy and decode overlap: one figure for both
  {
    uint64_t inside=n;
    if ((rc=applyWindow(ctx,n,scale,offset,unit,&inside)))
      return rc;
    return addSegment(ctx,inside,scale,offset,unit,kept,fmt,recLen);
  }
}

// ---- the reader pipeline: pread into a pinned ring on worker threads, H2D and decode behind it
namespace
{
int ensureRing(wb_ctx *ctx,uint64_t bufBytes)
// WB_READ_THREADS pinned buffers of at least bufBytes each (shared by the file reader and the file writer)
{
  if (ctx->readBufBytes<bufBytes)
  {
    for (int i=0;i<WB_READ_THREADS;i++)
    {
      if (ctx->readBuf[i])
        cudaFreeHost(ctx->readBuf[i]);
      ctx->readBuf[i]=nullptr;
    }
    ctx->readBufBytes=0;
    for (int i=0;i<WB_READ_THREADS;i++)
      if (cudaHostAlloc((void **)&ctx->readBuf[i],bufBytes,cudaHostAllocDefault)!=cudaSuccess)
        return fail(ctx,WB_ERR_NOMEM,"cannot pin %llu bytes for the file pipeline",(unsigned long long)bufBytes);
    ctx->readBufBytes=bufBytes;
  }
  for (int i=0;i<WB_READ_THREADS;i++)
    if (!ctx->evRead[i])
      CK(cudaEventCreateWithFlags(&ctx->evRead[i],cudaEventDisableTiming));
  return WB_OK;
}

struct ReadRing
{
  std::mutex m;
  std::condition_variable cv;
  int state[WB_READ_THREADS]={};   // 0 free, 1 filled, -1 read error
  bool stop=false;
};
}

extern "C" int wb_add_las_file(wb_ctx *ctx,const char *path,uint64_t pointOffset,uint64_t n,int fmt,int recLen,
                               const double scale[3],const double offset[3],double unit)
{
  if (!ctx || !path || !scale || !offset)
    return WB_ERR_ARG;
  cudaSetDevice(ctx->device);
  int rc=checkFormat(ctx,fmt,recLen);
  if (rc)
    return rc;
  if (ctx->phase>PH_LOADED)
    return fail(ctx,WB_ERR_STATE,"cloud already built: wb_clear first");
  int fd=open(path,O_RDONLY);
  if (fd<0)
    return fail(ctx,WB_ERR_ARG,"cannot open %s",path);
  struct FdGuard { int fd; ~FdGuard() { if (fd>=0) close(fd); } } fdGuard{fd};    // every early return closes the file
  // what a failure has to put back: the dropped-record count of the chunks already decoded (wb_build subtracts it)
  unsigned long long dropped0=0;
  CK(cudaMemcpyAsync(&dropped0,ctx->counters.p+2,sizeof(dropped0),cudaMemcpyDeviceToHost,ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  if ((rc=ensurePointArrays(ctx,ctx->n+n)))
    return rc;
  const int T=WB_READ_THREADS;
  const uint64_t chunkRecs=(uint64_t)1<<20;          // 1 Mi records (multiple of 16: chunks stay 16-byte aligned)
  const uint64_t chunkBytes=chunkRecs*recLen;
  const uint64_t nChunks=wb_div_up(n,chunkRecs);
  const uint64_t bufBytes=std::min<uint64_t>(chunkBytes,n*recLen)+64;
  if ((rc=ensureRing(ctx,bufBytes)))
    return rc;
  uint8_t *kept=nullptr;
  if (ctx->keepRecords && n)
    CK(cudaMalloc((void **)&kept,(size_t)(n*recLen+64)));
  else
    for (int b=0;b<2;b++)
      CK(ctx->staging[b].ensure(bufBytes));
  ReadRing ring;
  auto worker=[&](int t)
  { // chunk k lives in buffer k%T and is read by thread k%T
    for (uint64_t k=t;k<nChunks;k+=T)
    {
      {
        std::unique_lock<std::mutex> lk(ring.m);
        ring.cv.wait(lk,[&]{ return ring.state[t]==0 || ring.stop; });
        if (ring.stop)
          return;
      }
      uint64_t cnt=std::min(chunkRecs,n-k*chunkRecs),want=cnt*recLen,got=0;
      off_t pos=(off_t)(pointOffset+k*chunkBytes);
      while (got<want)
      {
        ssize_t r=pread(fd,ctx->readBuf[t]+got,want-got,pos+(off_t)got);
        if (r<=0)
          break;
        got+=(uint64_t)r;
      }
      {
        std::lock_guard<std::mutex> lk(ring.m);
        ring.state[t]=got==want?1:-1;
      }
      ring.cv.notify_all();
    }
  };
  std::vector<std::thread> threads;
  for (int t=0;t<T && (uint64_t)t<nChunks;t++)
    threads.emplace_back(worker,t);
  auto finish=[&](int code)
  {
    {
      std::lock_guard<std::mutex> lk(ring.m);
      ring.stop=true;
    }
    ring.cv.notify_all();
    for (auto &th:threads)
      th.join();
    if (code)
    { // leave the context as it was before the call: no half-read segment, no stale counts, no orphaned buffer
      cudaStreamSynchronize(ctx->stCopy);
      cudaStreamSynchronize(ctx->stLoad);
      cudaMemcpy(ctx->counters.p+2,&dropped0,sizeof(dropped0),cudaMemcpyHostToDevice);
      if (kept)
        cudaFree(kept);
    }
    return code;
  };
  cudaStream_t ld=ctx->stLoad;
  cudaEventRecord(ctx->evJoin,ctx->st);
  cudaStreamWaitEvent(ld,ctx->evJoin,0);
  cudaError_t ce=cudaEventRecord(ctx->evA,ld);
  int dropZeros=0,used[2]={0,0};
  for (uint64_t k=0;k<nChunks && ce==cudaSuccess;k++)
  {
    const int t=(int)(k%T),b=(int)(k&1);
    int stt;
    {
      std::unique_lock<std::mutex> lk(ring.m);
      ring.cv.wait(lk,[&]{ return ring.state[t]!=0; });
      stt=ring.state[t];
    }
    if (stt<0)
      return finish(fail(ctx,WB_ERR_ARG,"%s: short read (the header promises %llu points)",path,(unsigned long long)n));
    const uint64_t cnt=std::min(chunkRecs,n-k*chunkRecs);
    const uint8_t *src=ctx->readBuf[t];
    if (k==0)
      dropZeros=!ctx->keepZeroReturns && (fmt<6?(src[14]&7):(src[14]&15))!=0;   // threads.cpp:485-500: decided by record 0
    uint8_t *dst=kept?kept+k*chunkBytes:ctx->staging[b].p;
    if (used[b] && !kept)
      ce=cudaStreamWaitEvent(ctx->stCopy,ctx->evDec[b],0);
    if (ce==cudaSuccess) ce=cudaMemcpyAsync(dst,src,cnt*recLen,cudaMemcpyHostToDevice,ctx->stCopy);
    if (ce==cudaSuccess) ce=cudaEventRecord(ctx->evRead[t],ctx->stCopy);
    if (ce==cudaSuccess) ce=cudaStreamWaitEvent(ld,ctx->evRead[t],0);
    if (ce!=cudaSuccess)
      break;
    if ((rc=decodeDevice(ctx,dst,ctx->n+k*chunkRecs,cnt,fmt,recLen,dropZeros,ld)))
      return finish(rc);
    ce=cudaEventRecord(ctx->evDec[b],ld);
    used[b]=1;
    if (k>=1 && ce==cudaSuccess)
    { // the previous chunk's copy is done by now or soon: hand its buffer back to its reader
      const int tp=(int)((k-1)%T);
      ce=cudaEventSynchronize(ctx->evRead[tp]);
      {
        std::lock_guard<std::mutex> lk(ring.m);
        ring.state[tp]=0;
      }
      ring.cv.notify_all();
    }
  }
  if (ce!=cudaSuccess)
    return finish(fail(ctx,WB_ERR_CUDA,"%s",cudaGetErrorString(ce)));
  ce=cudaEventRecord(ctx->evB,ld);
  if (ce==cudaSuccess) ce=cudaStreamSynchronize(ctx->stCopy);
  if (ce==cudaSuccess) ce=cudaStreamSynchronize(ld);
  if (ce!=cudaSuccess)
    return finish(fail(ctx,WB_ERR_CUDA,"%s",cudaGetErrorString(ce)));
  finish(0);
  ctx->stats.ms_h2d+=elapsed(ctx->evA,ctx->evB);       // file read, copy and decode overlap: one figure
  if (kept)
    ctx->recBufs.push_back(kept);
  {
    uint64_t inside=n;
    if ((rc=applyWindow(ctx,n,scale,offset,unit,&inside)))
      return rc;
    return addSegment(ctx,inside,scale,offset,unit,kept,fmt,recLen);
  }
}

extern "C" int wb_add_points_device(wb_ctx *ctx,const int32_t *dx,const int32_t *dy,const int32_t *dz,const uint8_t *dc,
                                    uint64_t n,const double scale[3],const double offset[3],double unit)
{
  if (!ctx || (n && (!dx || !dy || !dz || !dc)) || !scale || !offset)
    return WB_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (ctx->phase>PH_LOADED)
    return fail(ctx,WB_ERR_STATE,"cloud already built: wb_clear first");
  int rc;
  if ((rc=ensurePointArrays(ctx,ctx->n+n)))
    return rc;
  if (n)
  {
    wb_copy_points_kernel<<<gridFor(n,256),256,0,ctx->st>>>(dx,dy,dz,dc,n,ctx->xi.p+ctx->n,ctx->yi.p+ctx->n,
                                                           ctx->zi.p+ctx->n,ctx->cls.p+ctx->n,ctx->ret.p+ctx->n);
    ctx->stats.kernel_launches++;
    KCHECK();
    CK(cudaStreamSynchronize(ctx->st));
  }
  return addSegment(ctx,n,scale,offset,unit);
}

extern "C" int wb_set_own_range(wb_ctx *ctx,uint64_t first,uint64_t end)
{
  if (!ctx || end<first)
    return WB_ERR_ARG;
  ctx->ownFirst=(uint32_t)first;
  ctx->ownEnd=end>0xffffffffull?0xffffffffu:(uint32_t)end;
  return WB_OK;
}

// ============================================================================ build

static int buildStore(wb_ctx *ctx,bool hilbert);

extern "C" int wb_build(wb_ctx *ctx)
{
  return buildStore(ctx,false);
}

static int buildStore(wb_ctx *ctx,bool hilbert)
// hilbert = a CLASSIFY-ONLY store (the second stage of wb_shard_run): the points sorted along the Hilbert curve
// classify walks anyway, no octree leaves; everything that needs the canonical order refuses such a store.
{
  if (!ctx)
    return WB_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (ctx->phase<PH_LOADED || ctx->n==0)
    return fail(ctx,WB_ERR_STATE,"no points loaded");
  if (ctx->n>=(1ull<<32)-64)
    return fail(ctx,WB_ERR_ARG,"more than 2^32 points per GPU are not supported");
  int rc=computeGeometry(ctx);
  if (rc)
    return rc;
  const uint64_t n=ctx->n;
  cudaStream_t st=ctx->st;
  CK(ctx->keyA.ensure(n)); CK(ctx->keyB.ensure(n)); CK(ctx->idxA.ensure(n)); CK(ctx->idxB.ensure(n));
  CK(ctx->sx.ensure(n)); CK(ctx->sy.ensure(n)); CK(ctx->sz.ensure(n)); CK(ctx->packed.ensure(n));
  CK(ctx->scr0.ensure(n+1)); CK(ctx->scr1.ensure(n+1));
  CK(ctx->leafDepth.ensure(n));
  CK(ctx->labelIn.ensure(n)); CK(ctx->labelSorted.ensure(n));
  CK(ctx->table.ensure(wb_div_up(n*3+16,WB_SORT_TILE)*256+256));
  CK(ctx->blockSums.ensure(wb_div_up(std::max<uint64_t>(n*3+16,ctx->table.cap),WB_SCAN_TILE)+1024));
  CK(cudaEventRecord(ctx->evA,st));
  // ---- keys
  for (size_t s=0;s<ctx->segs.size();s++)
  {
    const WbSegment &sg=ctx->segs[s];
    if (!sg.count)
      continue;
    wb_keygen_kernel<<<gridFor(sg.count,256),256,0,st>>>(ctx->xi.p,ctx->yi.p,ctx->zi.p,ctx->ret.p,sg.first,sg.count,sg,(int)s,
        ctx->geom.root_center[0],ctx->geom.root_center[1],ctx->geom.root_center[2],ctx->geom.root_side,hilbert?1:0,
        ctx->keyA.p,ctx->idxA.p,ctx->packed.p);
    ctx->stats.kernel_launches++;
  }
  KCHECK();
  // ---- sort (63 key bits -> 8 passes; Hilbert keys: 40 bits, the two special keys ~0 and WB_KEY_DUP have every
  //      higher bit set -> 48 bits order them behind the real ones, 6 passes)
  const int keyBits=hilbert?48:64;
  ctx->storeHilbert=hilbert;
  CK(cudaEventRecord(ctx->evC,st));
  bool inA=true;
  CK(wb_radix_sort((uint64_t *)ctx->keyA.p,ctx->idxA.p,(uint64_t *)ctx->keyB.p,ctx->idxB.p,n,0,keyBits,
                   ctx->table.p,ctx->table.cap,ctx->blockSums.p,ctx->blockSums.cap,st,&inA,&ctx->stats.kernel_launches));
  ctx->keys=inA?ctx->keyA.p:ctx->keyB.p;
  ctx->perm=inA?ctx->idxA.p:ctx->idxB.p;
  CK(cudaEventRecord(ctx->evD,st));
  // dropped records carry key ~0 and sit at the end
  unsigned long long dropped=0;
  CK(cudaMemcpyAsync(&dropped,ctx->counters.p+2,sizeof(dropped),cudaMemcpyDeviceToHost,st));
  CK(cudaStreamSynchronize(st));
  ctx->stats.ms_sort=elapsed(ctx->evC,ctx->evD);
  ctx->nValid=n-dropped;
  ctx->stats.n_dropped=dropped;
  ctx->stats.n_points=ctx->nValid;
  uint64_t nv=ctx->nValid;
  if (!nv)
    return fail(ctx,WB_ERR_STATE,"every record was dropped");
  // ---- canonical-order coordinates
  {
    WbSegments hs;
    hs.n=(int)ctx->segs.size();
    for (int i=0;i<hs.n;i++)
      hs.s[i]=ctx->segs[i];
    // only the entries in use: a host copy of more than a few KB is a copy-engine job and queues behind any bulk
    // upload another context of this device has pending (measured: the whole 147 KB table cost the pipelined e2e
    // its overlap, 499 -> 524 ms a step)
    CK(cudaMemcpyAsync(ctx->dsegs.p,&hs,offsetof(WbSegments,s)+sizeof(WbSegment)*(size_t)hs.n,cudaMemcpyHostToDevice,st));
    CK(cudaStreamSynchronize(st));
  }
  wb_gather_kernel<<<gridFor(nv,256),256,0,st>>>(ctx->perm,nv,ctx->packed.p,ctx->dsegs.p,ctx->sx.p,ctx->sy.p,ctx->sz.p);
  ctx->stats.kernel_launches++;
  KCHECK();
  // ---- identical locations: one point per XYZ stays in the store (octree.cpp:620-662)
  ctx->nDup=0;
  ctx->haveForced=false;
  {
    unsigned long long *cnt=ctx->counters.p+4;
    CK(cudaMemsetAsync(cnt,0,2*sizeof(unsigned long long),st));
    wb_dup_find_kernel<<<gridFor(nv,256),256,0,st>>>(ctx->keys,ctx->sx.p,ctx->sy.p,ctx->sz.p,nv,
                                                    ctx->scr0.p,ctx->scr1.p,cnt);
    ctx->stats.kernel_launches++;
    unsigned long long nDup=0;
    CK(cudaMemcpyAsync(&nDup,cnt,sizeof(nDup),cudaMemcpyDeviceToHost,st));
    CK(cudaStreamSynchronize(st));
    if (nDup)
    {
      if (nDup>=nv)
        return fail(ctx,WB_ERR_STATE,"internal: duplicate count");
      CK(ctx->dupIn.ensure(nDup)); CK(ctx->dupRep.ensure(nDup));
      for (int round=0;round<40;round++)
      {
        unsigned long long changed=0;
        CK(cudaMemsetAsync(cnt+1,0,sizeof(unsigned long long),st));
        wb_dup_jump_kernel<<<gridFor(nv,256),256,0,st>>>(ctx->scr0.p,ctx->scr1.p,nv,cnt+1);
        ctx->stats.kernel_launches++;
        CK(cudaMemcpyAsync(&changed,cnt+1,sizeof(changed),cudaMemcpyDeviceToHost,st));
        CK(cudaStreamSynchronize(st));
        if (!changed)
          break;
      }
      CK(cudaMemsetAsync(cnt+1,0,sizeof(unsigned long long),st));
      unsigned long long *curK=ctx->keys,*othK=(ctx->keys==ctx->keyA.p)?ctx->keyB.p:ctx->keyA.p;
      uint32_t *curP=ctx->perm,*othP=(ctx->perm==ctx->idxA.p)?ctx->idxB.p:ctx->idxA.p;
      // a halo point that holds the place of one of OUR records has to be classified here too
      ctx->haveForced=ctx->ownFirst!=0 || ctx->ownEnd!=0xffffffffu;
      if (ctx->haveForced)
      {
        CK(ctx->forcedIn.ensure(n));
        CK(cudaMemsetAsync(ctx->forcedIn.p,0,n,st));
      }
      wb_dup_mark_kernel<<<gridFor(nv,256),256,0,st>>>(ctx->scr0.p,ctx->scr1.p,curP,nv,curK,
                                                      ctx->dupIn.p,ctx->dupRep.p,cnt+1,
                                                      ctx->haveForced?ctx->forcedIn.p:nullptr,ctx->ownFirst,ctx->ownEnd);
      ctx->stats.kernel_launches++;
      KCHECK();
      // stable re-sort: survivors keep their order, duplicates go behind them (before the dropped records)
      bool inCur=true;
      CK(wb_radix_sort((uint64_t *)curK,curP,(uint64_t *)othK,othP,n,0,keyBits,ctx->table.p,ctx->table.cap,
                       ctx->blockSums.p,ctx->blockSums.cap,st,&inCur,&ctx->stats.kernel_launches));
      ctx->keys=inCur?curK:othK;
      ctx->perm=inCur?curP:othP;
      ctx->nDup=nDup;
      ctx->nValid=nv-nDup;
      ctx->stats.n_points=ctx->nValid;
      wb_gather_kernel<<<gridFor(ctx->nValid,256),256,0,st>>>(ctx->perm,ctx->nValid,ctx->packed.p,ctx->dsegs.p,
                                                             ctx->sx.p,ctx->sy.p,ctx->sz.p);
      ctx->stats.kernel_launches++;
      KCHECK();
    }
    ctx->stats.n_duplicates=ctx->nDup;
    nv=ctx->nValid;
  }
  // ---- leaves: level-synchronous top-down split
  CK(cudaEventRecord(ctx->evC,st));
  ctx->nLeaves=0;
  if (!hilbert)
  {
    uint64_t nodeCap=nv/256+64;
    CK(ctx->nodesA.ensure(nodeCap)); CK(ctx->nodesB.ensure(nodeCap));
    CK(cudaMemsetAsync(ctx->leafDepth.p,0,nv,st));
    CK(cudaMemsetAsync(ctx->counters.p+4,0,2*sizeof(unsigned long long),st));
    WbNode root;
    root.first=0;
    root.count=(uint32_t)nv;
    uint32_t nNodes=1;
    WbNode *cur=ctx->nodesA.p,*nxt=ctx->nodesB.p;
    CK(cudaMemcpyAsync(cur,&root,sizeof(root),cudaMemcpyHostToDevice,st));
    uint32_t *nNext=(uint32_t *)(ctx->counters.p+4),*nLeavesDev=(uint32_t *)(ctx->counters.p+5);
    for (int depth=0;depth<WB_LEVELS && nNodes;depth++)
    {
      CK(cudaMemsetAsync(nNext,0,sizeof(uint32_t),st));
      wb_split_kernel<<<gridFor((uint64_t)nNodes*8,128),128,0,st>>>(ctx->keys,cur,nNodes,depth,nxt,nNext,
                                                                    ctx->leafDepth.p,nLeavesDev);
      ctx->stats.kernel_launches++;
      CK(cudaMemcpyAsync(&nNodes,nNext,sizeof(uint32_t),cudaMemcpyDeviceToHost,st));
      CK(cudaStreamSynchronize(st));
      if (nNodes>nodeCap)
        return fail(ctx,WB_ERR_NOMEM,"internal: node list overflow");
      std::swap(cur,nxt);
    }
    KCHECK();
    CK(cudaMemcpyAsync(&ctx->nLeaves,nLeavesDev,sizeof(uint32_t),cudaMemcpyDeviceToHost,st));
    wb_leaf_flag_kernel<<<gridFor(nv,256),256,0,st>>>(ctx->leafDepth.p,nv,ctx->scr0.p);
    CK(wb_exclusive_scan(ctx->scr0.p,ctx->scr1.p,nv,ctx->blockSums.p,ctx->blockSums.cap,st,&ctx->stats.kernel_launches));
    CK(cudaStreamSynchronize(st));
    CK(ctx->leaves.ensure(ctx->nLeaves+1));
    wb_leaf_emit_kernel<<<gridFor(nv,256),256,0,st>>>(ctx->leafDepth.p,ctx->scr1.p,nv,ctx->leaves.p);
    wb_leaf_finish_kernel<<<gridFor((uint64_t)ctx->nLeaves*32,256),256,0,st>>>(ctx->leaves.p,ctx->nLeaves,nv,ctx->sz.p);
    ctx->stats.kernel_launches+=3;
    KCHECK();
  }
  CK(cudaEventRecord(ctx->evD,st));
  // ---- bucket chunks and the 32-ary hierarchy over them
  {
    ctx->nChunks=(uint32_t)wb_div_up(nv,32);
    ctx->hLevelOff.clear();
    ctx->hLevelCnt.clear();
    uint64_t total=0;
    uint32_t c=ctx->nChunks;
    while (true)
    {
      ctx->hLevelOff.push_back((uint32_t)total);
      ctx->hLevelCnt.push_back(c);
      total+=c;
      if (c<=32)
        break;
      c=(uint32_t)wb_div_up(c,32);
    }
    ctx->nLevels=(int)ctx->hLevelCnt.size();
    CK(ctx->bounds.ensure(total));
    CK(ctx->levelOff.ensure(16)); CK(ctx->levelCnt.ensure(16));
    CK(cudaMemcpyAsync(ctx->levelOff.p,ctx->hLevelOff.data(),ctx->nLevels*sizeof(uint32_t),cudaMemcpyHostToDevice,st));
    CK(cudaMemcpyAsync(ctx->levelCnt.p,ctx->hLevelCnt.data(),ctx->nLevels*sizeof(uint32_t),cudaMemcpyHostToDevice,st));
    wb_chunk_bounds_kernel<<<gridFor((uint64_t)ctx->nChunks*32,256),256,0,st>>>(ctx->sx.p,ctx->sy.p,ctx->sz.p,nv,
                                                                              ctx->bounds.p,ctx->nChunks);
    ctx->stats.kernel_launches++;
    for (int l=1;l<ctx->nLevels;l++)
    {
      wb_node_bounds_kernel<<<gridFor((uint64_t)ctx->hLevelCnt[l]*32,256),256,0,st>>>(
          ctx->bounds.p+ctx->hLevelOff[l-1],ctx->hLevelCnt[l-1],ctx->bounds.p+ctx->hLevelOff[l],ctx->hLevelCnt[l]);
      ctx->stats.kernel_launches++;
    }
    KCHECK();
  }
  CK(cudaEventRecord(ctx->evB,st));
  CK(cudaStreamSynchronize(st));
  ctx->stats.ms_leaves=elapsed(ctx->evC,ctx->evD);
  ctx->stats.ms_hier=elapsed(ctx->evD,ctx->evB);
  ctx->stats.ms_build=elapsed(ctx->evA,ctx->evB);
  ctx->stats.n_leaves=ctx->nLeaves;
  ctx->phase=PH_BUILT;
  return WB_OK;
}

extern "C" int wb_num_leaves(wb_ctx *ctx,uint64_t *n)
{
  if (!ctx || !n)
    return WB_ERR_ARG;
  if (ctx->phase<PH_BUILT)
    return fail(ctx,WB_ERR_STATE,"not built");
  if (ctx->storeHilbert)
    return fail(ctx,WB_ERR_STATE,"classify-only store (second stage of wb_shard_run): it has no canonical order or leaves");
  *n=ctx->nLeaves;
  return WB_OK;
}

extern "C" int wb_get_leaves(wb_ctx *ctx,wb_leaf *out,uint64_t cap)
{
  if (!ctx || !out)
    return WB_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (ctx->phase<PH_BUILT)
    return fail(ctx,WB_ERR_STATE,"not built");
  if (ctx->storeHilbert)
    return fail(ctx,WB_ERR_STATE,"classify-only store (second stage of wb_shard_run): it has no canonical order or leaves");
  if (cap<ctx->nLeaves)
    return fail(ctx,WB_ERR_ARG,"leaf buffer too small");
  std::vector<WbLeafDev> h(ctx->nLeaves);
  std::vector<unsigned long long> firstKey(ctx->nLeaves);
  CK(cudaMemcpy(h.data(),ctx->leaves.p,sizeof(WbLeafDev)*ctx->nLeaves,cudaMemcpyDeviceToHost));
  // key of each leaf's first point -> cube
  {
    DevBuf<unsigned long long> tmp;
    CK(tmp.ensure(ctx->nLeaves+1));
    wb_leaf_keys_kernel<<<gridFor(ctx->nLeaves,256),256,0,ctx->st>>>(ctx->leaves.p,ctx->nLeaves,ctx->keys,tmp.p);
    ctx->stats.kernel_launches++;
    KCHECK();
    CK(cudaMemcpyAsync(firstKey.data(),tmp.p,sizeof(unsigned long long)*ctx->nLeaves,cudaMemcpyDeviceToHost,ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    tmp.release();
  }
  for (uint32_t i=0;i<ctx->nLeaves;i++)
  {
    double c[3],half;
    wbhost::leafCube(firstKey[i],h[i].depth,ctx->geom.root_center,ctx->geom.root_side,c,&half);
    out[i].first=h[i].first;
    out[i].count=h[i].count;
    out[i].depth=h[i].depth;
    out[i].cx=c[0]; out[i].cy=c[1]; out[i].cz=c[2];
    out[i].half=half;
    out[i].low=h[i].low;
    out[i].high=h[i].high;
  }
  return WB_OK;
}

extern "C" int wb_get_order(wb_ctx *ctx,uint32_t *order,uint64_t *keys)
{
  if (!ctx)
    return WB_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (ctx->phase<PH_BUILT)
    return fail(ctx,WB_ERR_STATE,"not built");
  if (ctx->storeHilbert)
    return fail(ctx,WB_ERR_STATE,"classify-only store (second stage of wb_shard_run): it has no canonical order or leaves");
  if (order)
    CK(cudaMemcpy(order,ctx->perm,sizeof(uint32_t)*ctx->nValid,cudaMemcpyDeviceToHost));
  if (keys)
    CK(cudaMemcpy(keys,ctx->keys,sizeof(uint64_t)*ctx->nValid,cudaMemcpyDeviceToHost));
  return WB_OK;
}

extern "C" int wb_get_points_sorted(wb_ctx *ctx,double *x,double *y,double *z)
{
  if (!ctx)
    return WB_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (ctx->phase<PH_BUILT)
    return fail(ctx,WB_ERR_STATE,"not built");
  if (ctx->storeHilbert)
    return fail(ctx,WB_ERR_STATE,"classify-only store (second stage of wb_shard_run): it has no canonical order or leaves");
  if (x) CK(cudaMemcpy(x,ctx->sx.p,sizeof(double)*ctx->nValid,cudaMemcpyDeviceToHost));
  if (y) CK(cudaMemcpy(y,ctx->sy.p,sizeof(double)*ctx->nValid,cudaMemcpyDeviceToHost));
  if (z) CK(cudaMemcpy(z,ctx->sz.p,sizeof(double)*ctx->nValid,cudaMemcpyDeviceToHost));
  return WB_OK;
}

extern "C" int wb_get_decoded(wb_ctx *ctx,int32_t *x,int32_t *y,int32_t *z,uint8_t *cls)
{
  if (!ctx)
    return WB_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (ctx->phase<PH_LOADED)
    return fail(ctx,WB_ERR_STATE,"nothing loaded");
  if (x) CK(cudaMemcpy(x,ctx->xi.p,sizeof(int)*ctx->n,cudaMemcpyDeviceToHost));
  if (y) CK(cudaMemcpy(y,ctx->yi.p,sizeof(int)*ctx->n,cudaMemcpyDeviceToHost));
  if (z) CK(cudaMemcpy(z,ctx->zi.p,sizeof(int)*ctx->n,cudaMemcpyDeviceToHost));
  if (cls) CK(cudaMemcpy(cls,ctx->cls.p,ctx->n,cudaMemcpyDeviceToHost));
  return WB_OK;
}

// ============================================================================ scan / postscan

extern "C" int wb_scan(wb_ctx *ctx)
{
  if (!ctx)
    return WB_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (ctx->phase<PH_BUILT)
    return fail(ctx,WB_ERR_STATE,"not built");
  if (ctx->storeHilbert)
    return fail(ctx,WB_ERR_STATE,"classify-only store (second stage of wb_shard_run): it has no canonical order or leaves");
  const uint64_t nv=ctx->nValid;
  const uint32_t T=ctx->nTiles;
  cudaStream_t st=ctx->st;
  CK(ctx->tilesOf.ensure(nv)); CK(ctx->winner.ensure(nv));
  {
    int rc=ensureTileArrays(ctx);
    if (rc)
      return rc;
  }
  CK(cudaEventRecord(ctx->evA,st));
  wb_member_count_kernel<<<gridFor(nv,128),128,0,st>>>(ctx->sx.p,ctx->sy.p,nv,ctx->snake,ctx->scr0.p,ctx->tilesOf.p,ctx->winner.p);
  ctx->stats.kernel_launches++;
  KCHECK();
  CK(cudaMemsetAsync(ctx->scr0.p+nv,0,sizeof(uint32_t),st));
  CK(wb_exclusive_scan(ctx->scr0.p,ctx->scr1.p,nv+1,ctx->blockSums.p,ctx->blockSums.cap,st,&ctx->stats.kernel_launches));
  uint32_t m=0;
  CK(cudaMemcpyAsync(&m,ctx->scr1.p+nv,sizeof(uint32_t),cudaMemcpyDeviceToHost,st));
  CK(cudaStreamSynchronize(st));
  ctx->nPairs=m;
  CK(ctx->pairKeyA.ensure(m+1)); CK(ctx->pairKeyB.ensure(m+1)); CK(ctx->pairValA.ensure(m+1)); CK(ctx->pairValB.ensure(m+1));
  CK(ctx->table.ensure(wb_div_up((uint64_t)m+16,WB_SORT_TILE)*256+256));
  CK(ctx->blockSums.ensure(wb_div_up(ctx->table.cap,WB_SCAN_TILE)+1024));
  wb_member_fill_kernel<<<gridFor(nv,256),256,0,st>>>(ctx->scr0.p,ctx->scr1.p,ctx->tilesOf.p,nv,ctx->pairKeyA.p,ctx->pairValA.p);
  ctx->stats.kernel_launches++;
  KCHECK();
  int bits=1;
  while (bits<32 && (1ull<<bits)<T)
    bits++;
  bits=(bits+7)/8*8;
  bool inA=true;
  CK(wb_radix_sort(ctx->pairKeyA.p,ctx->pairValA.p,ctx->pairKeyB.p,ctx->pairValB.p,m,0,bits,
                   ctx->table.p,ctx->table.cap,ctx->blockSums.p,ctx->blockSums.cap,st,&inA,&ctx->stats.kernel_launches));
  ctx->pairKeys=inA?ctx->pairKeyA.p:ctx->pairKeyB.p;
  ctx->pairVals=inA?ctx->pairValA.p:ctx->pairValB.p;
  CK(cudaMemsetAsync(ctx->tNPoints.p,0,sizeof(int)*T,st));
  CK(cudaMemsetAsync(ctx->counters.p+3,0,sizeof(unsigned long long),st));
  CK(ctx->tileList.ensure(std::min<uint64_t>(T,(uint64_t)m)+1));
  if (m)
    wb_segment_kernel<<<gridFor(m,256),256,0,st>>>(ctx->pairKeys,m,ctx->tStart.p,ctx->tCount.p,
                                                   ctx->tileList.p,ctx->counters.p+3);
  CK(cudaEventRecord(ctx->evC,st));
  unsigned long long nList=0;
  CK(cudaMemcpyAsync(&nList,ctx->counters.p+3,sizeof(nList),cudaMemcpyDeviceToHost,st));
  CK(cudaStreamSynchronize(st));
  if (nList)
  {
    const int bundle=wb_scan_bundle(nList,ctx->nPairs);
    wb_scan_kernel<<<gridFor(nList,(unsigned)(WB_SCAN_WARPS*bundle)),WB_SCAN_WARPS*32,0,st>>>(ctx->tileList.p,(uint32_t)nList,ctx->tStart.p,ctx->tCount.p,
                                              ctx->pairVals,ctx->sx.p,ctx->sy.p,ctx->sz.p,
                                              ctx->snake,ctx->prm.minHyp,bundle,ctx->tNPoints.p,ctx->tTree.p,ctx->tDensity.p,
                                              ctx->tHyp.p,ctx->tHeight.p);
  }
  ctx->stats.kernel_launches+=2;
  KCHECK();
  CK(cudaEventRecord(ctx->evB,st));
  unsigned long long ne=0;
  CK(cudaMemcpyAsync(&ne,ctx->counters.p+3,sizeof(ne),cudaMemcpyDeviceToHost,st));
  CK(cudaStreamSynchronize(st));
  ctx->stats.ms_pairs=elapsed(ctx->evA,ctx->evC);
  ctx->stats.ms_scan=elapsed(ctx->evA,ctx->evB);
  ctx->stats.n_tiles_nonempty=ne;
  ctx->stats.n_memberships=m;
  ctx->phase=PH_SCANNED;
  return WB_OK;
}

static int ensureTileArrays(wb_ctx *ctx)
{
  const uint32_t T=ctx->nTiles;
  CK(ctx->tStart.ensure(T)); CK(ctx->tCount.ensure(T)); CK(ctx->tNPoints.ensure(T)); CK(ctx->tTree.ensure(T));
  CK(ctx->tDensity.ensure(T)); CK(ctx->tHyp.ensure(T)); CK(ctx->tHeight.ensure(T));
  return WB_OK;
}

extern "C" int wb_assign(wb_ctx *ctx)
{
  if (!ctx)
    return WB_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (ctx->phase<PH_BUILT)
    return fail(ctx,WB_ERR_STATE,"not built");
  const uint64_t nv=ctx->nValid;
  CK(ctx->tilesOf.ensure(nv)); CK(ctx->winner.ensure(nv));
  wb_member_count_kernel<<<gridFor(nv,128),128,0,ctx->st>>>(ctx->sx.p,ctx->sy.p,nv,ctx->snake,ctx->scr0.p,ctx->tilesOf.p,ctx->winner.p);
  ctx->stats.kernel_launches++;
  KCHECK();
  CK(cudaStreamSynchronize(ctx->st));
  return WB_OK;
}

extern "C" int wb_max_hyperboloid_size(wb_ctx *ctx,double *out)
{
  if (!ctx || !out)
    return WB_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (ctx->phase<PH_SCANNED)
    return fail(ctx,WB_ERR_STATE,"not scanned");
  CK(cudaMemsetAsync(ctx->counters.p+7,0,sizeof(unsigned long long),ctx->st));
  if (ctx->stats.n_tiles_nonempty)
    wb_max_hyp_list_kernel<<<148*4,256,0,ctx->st>>>(ctx->tileList.p,(uint32_t)ctx->stats.n_tiles_nonempty,ctx->tHyp.p,ctx->snake,
                                                   -INFINITY,INFINITY,ctx->counters.p+7);
  ctx->stats.kernel_launches++;
  KCHECK();
  unsigned long long bits=0;
  CK(cudaMemcpyAsync(&bits,ctx->counters.p+7,sizeof(bits),cudaMemcpyDeviceToHost,ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  memcpy(out,&bits,sizeof(double));
  return WB_OK;
}

extern "C" int wb_postscan(wb_ctx *ctx)
{
  if (!ctx)
    return WB_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (ctx->phase!=PH_SCANNED)
    return fail(ctx,WB_ERR_STATE,"postscan follows scan (and runs once)");
  cudaStream_t st=ctx->st;
  CK(cudaEventRecord(ctx->evA,st));
  {
    const int init[4]={INT_MAX,INT_MAX,INT_MIN,INT_MIN};
    int ext[4];
    const uint32_t nList=(uint32_t)ctx->stats.n_tiles_nonempty;
    CK(ctx->tileExt.ensure(4));
    CK(cudaMemcpyAsync(ctx->tileExt.p,init,sizeof(init),cudaMemcpyHostToDevice,st));
    if (nList)
      wb_tile_extent_list_kernel<<<gridFor(nList,256),256,0,st>>>(ctx->tileList.p,nList,ctx->snake,-INFINITY,INFINITY,ctx->tileExt.p);
    CK(cudaMemcpyAsync(ext,ctx->tileExt.p,sizeof(ext),cudaMemcpyDeviceToHost,st));
    CK(cudaStreamSynchronize(st));
    uint64_t cells=1;
    if (ext[0]<=ext[2])
      cells=(uint64_t)((long long)ext[2]-ext[0]+1)*(uint64_t)((long long)ext[3]-ext[1]+1);
    CK(ctx->tileGrid.ensure(cells));
    CK(cudaMemsetAsync(ctx->tileGrid.p,0,cells,st));
    if (nList)
    {
      wb_tile_grid_list_kernel<<<gridFor(nList,256),256,0,st>>>(ctx->tileList.p,nList,ctx->tTree.p,ctx->snake,-INFINITY,INFINITY,
                                                               ctx->tileExt.p,ctx->tileGrid.p);
      wb_postscan_list_kernel<<<gridFor(nList,128),128,0,st>>>(ctx->tileList.p,nList,ctx->tTree.p,ctx->snake,
                                                              ctx->tileExt.p,ctx->tileGrid.p,ctx->tHyp.p);
    }
    ctx->stats.kernel_launches+=3;
  }
  KCHECK();
  CK(cudaEventRecord(ctx->evB,st));
  CK(cudaStreamSynchronize(st));
  ctx->stats.ms_postscan=elapsed(ctx->evA,ctx->evB);
  ctx->phase=PH_POSTSCANNED;
  return WB_OK;
}

extern "C" int wb_num_tiles(wb_ctx *ctx,uint64_t *n)
{
  if (!ctx || !n)
    return WB_ERR_ARG;
  if (ctx->phase<PH_SCANNED)
    return fail(ctx,WB_ERR_STATE,"not scanned");
  *n=ctx->stats.n_tiles_nonempty;
  return WB_OK;
}

extern "C" int wb_get_tiles(wb_ctx *ctx,wb_tile *out,uint64_t cap)
{
  if (!ctx || !out)
    return WB_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (ctx->phase<PH_SCANNED)
    return fail(ctx,WB_ERR_STATE,"not scanned");
  if (cap<ctx->stats.n_tiles_nonempty)
    return fail(ctx,WB_ERR_ARG,"tile buffer too small");
  const uint32_t T=ctx->nTiles;
  std::vector<int> np(T);
  std::vector<uint8_t> tr(T);
  std::vector<double> de(T),hy(T),he(T);
  CK(cudaMemcpy(np.data(),ctx->tNPoints.p,sizeof(int)*T,cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(tr.data(),ctx->tTree.p,T,cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(de.data(),ctx->tDensity.p,sizeof(double)*T,cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hy.data(),ctx->tHyp.p,sizeof(double)*T,cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(he.data(),ctx->tHeight.p,sizeof(double)*T,cudaMemcpyDeviceToHost));
  uint64_t k=0;
  for (uint32_t t=0;t<T;t++)
    if (np[t])
    {
      if (k>=cap)
        break;
      wb_tile &o=out[k++];
      o.n=(int32_t)((long long)t+ctx->snake.lo);
      // Eisenstein address: same integer recipe as the device (toFlowsnake)
      {
        int dig[11],ori=0;
        long long v=(long long)o.n+1235829214LL;
        for (int i=0;i<11;i++) { dig[i]=(int)(v%7); v/=7; }
        for (int i=10;i>=0;i--) { int tt=wbhost::kFwdTable[ori][dig[i]]; ori=tt>>4; dig[i]=tt&7; }
        int x=0,y=0,px=1,py=0;
        for (int i=0;i<11;i++)
        {
          int d=dig[i]-3,dy=(d+4)/3-1,dx=d-2*dy;
          x+=dx*px-dy*py;
          y+=dx*py+dy*px-dy*py;
          int nx=2*px+py,ny=3*py-px;
          px=nx; py=ny;
        }
        o.ex=x; o.ey=y;
      }
      o.nPoints=np[t];
      o.treeFlags=tr[t];
      o.pad_=0;
      o.density=de[t];
      o.hyperboloidSize=hy[t];
      o.height=he[t];
    }
  return WB_OK;
}

extern "C" int wb_set_tiles(wb_ctx *ctx,const wb_tile *tiles,uint64_t n)
{
  if (!ctx || (!tiles && n))
    return WB_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (ctx->phase<PH_SCANNED)
    return fail(ctx,WB_ERR_STATE,"scan first (the tile table must exist)");
  const uint32_t T=ctx->nTiles;
  std::vector<double> hy(T);
  CK(cudaMemcpy(hy.data(),ctx->tHyp.p,sizeof(double)*T,cudaMemcpyDeviceToHost));
  for (uint64_t i=0;i<n;i++)
  {
    long long t=(long long)tiles[i].n-ctx->snake.lo;
    if (t<0 || t>=(long long)T)
      return fail(ctx,WB_ERR_ARG,"tile %d outside the flowsnake range",tiles[i].n);
    hy[t]=tiles[i].hyperboloidSize;
  }
  CK(cudaMemcpy(ctx->tHyp.p,hy.data(),sizeof(double)*T,cudaMemcpyHostToDevice));
  if (ctx->phase<PH_POSTSCANNED)
    ctx->phase=PH_POSTSCANNED;
  return WB_OK;
}

// ============================================================================ classify

extern "C" int wb_classify(wb_ctx *ctx)
{
  if (!ctx)
    return WB_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (ctx->phase<PH_POSTSCANNED)
    return fail(ctx,WB_ERR_STATE,"classify follows postscan");
  const uint64_t nv=ctx->nValid;
  cudaStream_t st=ctx->st;
  CK(cudaEventRecord(ctx->evA,st));
  CK(cudaMemsetAsync(ctx->counters.p,0,2*sizeof(unsigned long long),st));
  CK(cudaMemsetAsync(ctx->counters.p+6,0,18*sizeof(unsigned long long),st));
  wb_init_labels_kernel<<<gridFor(ctx->n,256),256,0,st>>>(ctx->cls.p,ctx->n,ctx->labelIn.p);
  const double *qx=ctx->sx.p,*qy=ctx->sy.p,*qz=ctx->sz.p;
  const uint32_t *qwinner=ctx->winner.p,*qperm=ctx->perm,*hord=nullptr;
  const WbBound *qbounds=ctx->bounds.p;
  uint8_t *qlabel=ctx->labelSorted.p;
#if WB_CL_HILBERT
  if (nv && !ctx->storeHilbert)
  {
    // the store along a Hilbert curve over xy (see wb_kernels.cuh, "classify order")
    CK(ctx->hKeyA.ensure(nv)); CK(ctx->hKeyB.ensure(nv)); CK(ctx->hIdxA.ensure(nv)); CK(ctx->hIdxB.ensure(nv));
    CK(ctx->hx.ensure(nv)); CK(ctx->hy.ensure(nv)); CK(ctx->hz.ensure(nv)); CK(ctx->hWinner.ensure(nv)); CK(ctx->hPerm.ensure(nv));
    CK(ctx->hLabel.ensure(nv));
    CK(ctx->hBounds.ensure(ctx->hLevelOff.back()+ctx->hLevelCnt.back()));
    const wb_geometry &g=ctx->geom;
    const double x0=g.root_center[0]-g.root_side,y0=g.root_center[1]-g.root_side;
    const double cells=(double)(1u<<WB_HILBERT_BITS)/(2*g.root_side);
    wb_hilbert_key_kernel<<<gridFor(nv,256),256,0,st>>>(ctx->sx.p,ctx->sy.p,nv,x0,y0,cells,ctx->hKeyA.p,ctx->hIdxA.p);
    ctx->stats.kernel_launches++;
    bool inA=true;
    CK(wb_radix_sort((uint64_t *)ctx->hKeyA.p,ctx->hIdxA.p,(uint64_t *)ctx->hKeyB.p,ctx->hIdxB.p,nv,0,2*WB_HILBERT_BITS,
                     ctx->table.p,ctx->table.cap,ctx->blockSums.p,ctx->blockSums.cap,st,&inA,&ctx->stats.kernel_launches));
    hord=inA?ctx->hIdxA.p:ctx->hIdxB.p;
    wb_classify_gather_kernel<<<gridFor(nv,256),256,0,st>>>(hord,nv,ctx->sx.p,ctx->sy.p,ctx->sz.p,ctx->winner.p,ctx->perm,
                                                           ctx->hx.p,ctx->hy.p,ctx->hz.p,ctx->hWinner.p,ctx->hPerm.p);
    wb_chunk_bounds_kernel<<<gridFor((uint64_t)ctx->nChunks*32,256),256,0,st>>>(ctx->hx.p,ctx->hy.p,ctx->hz.p,nv,
                                                                              ctx->hBounds.p,ctx->nChunks);
    ctx->stats.kernel_launches+=2;
    for (int l=1;l<ctx->nLevels;l++)
    {
      wb_node_bounds_kernel<<<gridFor((uint64_t)ctx->hLevelCnt[l]*32,256),256,0,st>>>(
          ctx->hBounds.p+ctx->hLevelOff[l-1],ctx->hLevelCnt[l-1],ctx->hBounds.p+ctx->hLevelOff[l],ctx->hLevelCnt[l]);
      ctx->stats.kernel_launches++;
    }
    KCHECK();
    qx=ctx->hx.p; qy=ctx->hy.p; qz=ctx->hz.p;
    qwinner=ctx->hWinner.p; qperm=ctx->hPerm.p;
    qbounds=ctx->hBounds.p;
    qlabel=ctx->hLabel.p;
  }
#endif
  CK(cudaEventRecord(ctx->evC,st));
  CK(ctx->wedgeBuf.ensure(nv));
  CK(ctx->chunkPending.ensure(ctx->nChunks));
  CK(cudaMemsetAsync(ctx->chunkPending.p,0,ctx->nChunks,st));
  wb_classify_kernel<1><<<gridFor(ctx->nChunks,WB_CL_WARPS),WB_CL_WARPS*32,0,st>>>(
      qx,qy,qz,nv,ctx->nChunks,qbounds,ctx->levelOff.p,ctx->levelCnt.p,ctx->nLevels,
      qwinner,ctx->tHyp.p,ctx->prm.maxSlope,ctx->prm.thickness,ctx->cls.p,qperm,
      ctx->ownFirst,ctx->ownEnd,ctx->haveForced?ctx->forcedIn.p:nullptr,qlabel,ctx->counters.p,ctx->wedgeBuf.p,ctx->chunkPending.p
#if WB_CL_COMPACT2
      ,nullptr,0u
#endif
      );
#if WB_CL_COMPACT2
  {
    // the queries that need the exact walk, gathered in canonical order into full warps
    wb_pending_flag_kernel<<<gridFor(nv,256),256,0,st>>>(ctx->wedgeBuf.p,nv,ctx->scr0.p);
    CK(cudaMemsetAsync(ctx->scr0.p+nv,0,sizeof(uint32_t),st));
    CK(wb_exclusive_scan(ctx->scr0.p,ctx->scr1.p,nv+1,ctx->blockSums.p,ctx->blockSums.cap,st,&ctx->stats.kernel_launches));
    uint32_t nPending=0;
    CK(cudaMemcpyAsync(&nPending,ctx->scr1.p+nv,sizeof(uint32_t),cudaMemcpyDeviceToHost,st));
    CK(cudaStreamSynchronize(st));
    CK(ctx->pendingList.ensure((uint64_t)nPending+1));
    if (nPending)
    {
      wb_pending_scatter_kernel<<<gridFor(nv,256),256,0,st>>>(ctx->scr0.p,ctx->scr1.p,nv,ctx->pendingList.p);
      wb_classify_kernel<2><<<gridFor(wb_div_up(nPending,32),WB_CL_WARPS),WB_CL_WARPS*32,0,st>>>(
          qx,qy,qz,nv,ctx->nChunks,qbounds,ctx->levelOff.p,ctx->levelCnt.p,ctx->nLevels,
          qwinner,ctx->tHyp.p,ctx->prm.maxSlope,ctx->prm.thickness,ctx->cls.p,qperm,
          ctx->ownFirst,ctx->ownEnd,ctx->haveForced?ctx->forcedIn.p:nullptr,qlabel,ctx->counters.p,ctx->wedgeBuf.p,ctx->chunkPending.p,
          ctx->pendingList.p,nPending);
    }
    ctx->stats.kernel_launches+=2;
  }
#else
  wb_classify_kernel<2><<<gridFor(ctx->nChunks,WB_CL_WARPS),WB_CL_WARPS*32,0,st>>>(
      qx,qy,qz,nv,ctx->nChunks,qbounds,ctx->levelOff.p,ctx->levelCnt.p,ctx->nLevels,
      qwinner,ctx->tHyp.p,ctx->prm.maxSlope,ctx->prm.thickness,ctx->cls.p,qperm,
      ctx->ownFirst,ctx->ownEnd,ctx->haveForced?ctx->forcedIn.p:nullptr,qlabel,ctx->counters.p,ctx->wedgeBuf.p,ctx->chunkPending.p);
#endif
  CK(cudaEventRecord(ctx->evD,st));
  if (hord)
  {
    wb_classify_scatter_kernel<<<gridFor(nv,256),256,0,st>>>(hord,ctx->hLabel.p,nv,ctx->labelSorted.p);
    ctx->stats.kernel_launches++;
  }
  wb_scatter_labels_kernel<<<gridFor(nv,256),256,0,st>>>(ctx->labelSorted.p,ctx->perm,nv,ctx->labelIn.p);
  ctx->stats.kernel_launches+=4;
  if (ctx->nDup)
  { // records lost to an identical location take the class of the point that stayed in the store
    wb_dup_labels_kernel<<<gridFor(ctx->nDup,256),256,0,st>>>(ctx->dupIn.p,ctx->dupRep.p,ctx->nDup,ctx->labelIn.p);
    ctx->stats.kernel_launches++;
  }
  KCHECK();
  CK(cudaEventRecord(ctx->evB,st));
  unsigned long long c[24]={0};
  CK(cudaMemcpyAsync(c,ctx->counters.p,sizeof(c),cudaMemcpyDeviceToHost,st));
  CK(cudaStreamSynchronize(st));
  ctx->stats.n_second_walk=c[6];
  ctx->stats.cl_nodes=c[8];
  ctx->stats.cl_chunks=c[9];
  ctx->stats.cl_pairs=c[10];
  ctx->stats.cl_nodes2=c[11];
  ctx->stats.cl_chunks2=c[12];
  ctx->stats.cl_pairs2=c[13];
  ctx->stats.cl_warps2=c[14];
  ctx->stats.ms_classify=elapsed(ctx->evA,ctx->evB);
  ctx->stats.ms_classify_kernel=elapsed(ctx->evC,ctx->evD);
  ctx->stats.ms_classify_order=elapsed(ctx->evA,ctx->evC);
  ctx->stats.n_margin=c[0];
  ctx->stats.n_untiled=c[1];
  ctx->phase=PH_CLASSIFIED;
  return WB_OK;
}

extern "C" int wb_get_labels(wb_ctx *ctx,uint8_t *labels)
{
  if (!ctx || !labels)
    return WB_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (ctx->phase<PH_CLASSIFIED)
    return fail(ctx,WB_ERR_STATE,"not classified");
  CK(cudaEventRecord(ctx->evA,ctx->st));
  CK(cudaMemcpyAsync(labels,ctx->labelIn.p,ctx->n,cudaMemcpyDeviceToHost,ctx->st));
  CK(cudaEventRecord(ctx->evB,ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  ctx->stats.ms_d2h=elapsed(ctx->evA,ctx->evB);
  return WB_OK;
}

extern "C" int wb_count_classes(wb_ctx *ctx,uint64_t counts[256])
{
  if (!ctx || !counts)
    return WB_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (ctx->phase<PH_CLASSIFIED)
    return fail(ctx,WB_ERR_STATE,"not classified");
  DevBuf<unsigned long long> d;
  CK(d.ensure(256));
  CK(cudaMemsetAsync(d.p,0,256*sizeof(unsigned long long),ctx->st));
  wb_count_classes_kernel<<<148*4,256,0,ctx->st>>>(ctx->labelIn.p,ctx->ret.p,ctx->n,d.p);
  ctx->stats.kernel_launches++;
  KCHECK();
  CK(cudaMemcpyAsync(counts,d.p,256*sizeof(unsigned long long),cudaMemcpyDeviceToHost,ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  d.release();
  return WB_OK;
}

// ============================================================================ store queries (OctStore)

namespace
{
int checkShapes(wb_ctx *ctx,const wb_shape *sh,uint64_t n)
{
  for (uint64_t i=0;i<n;i++)
    if (sh[i].type<WB_SHAPE_SPHERE || sh[i].type>WB_SHAPE_COLUMN)
      return fail(ctx,WB_ERR_ARG,"shape %llu: unknown type %d",(unsigned long long)i,sh[i].type);
  return WB_OK;
}
}

extern "C" int wb_query_batch(wb_ctx *ctx,const wb_shape *shapes,uint64_t n,uint64_t *count,double *lo,double *hi)
{
  if (!ctx || (n && !shapes))
    return WB_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (ctx->phase<PH_BUILT)
    return fail(ctx,WB_ERR_STATE,"not built");
  if (!n)
    return WB_OK;
  int rc=checkShapes(ctx,shapes,n);
  if (rc)
    return rc;
  static_assert(sizeof(wb_shape)==sizeof(WbShapeDev),"shape layout");
  DevBuf<WbShapeDev> ds;
  DevBuf<unsigned long long> dc;
  DevBuf<double> dl,dh;
  cudaStream_t st=ctx->st;
  CK(ds.ensure(n)); CK(dc.ensure(n)); CK(dl.ensure(n)); CK(dh.ensure(n));
  CK(cudaMemcpyAsync(ds.p,shapes,sizeof(wb_shape)*n,cudaMemcpyHostToDevice,st));
  wb_query_kernel<<<(unsigned)wb_div_up(n,WB_Q_WARPS),WB_Q_WARPS*32,0,st>>>(ds.p,n,ctx->sx.p,ctx->sy.p,ctx->sz.p,ctx->nValid,
      ctx->bounds.p,ctx->levelOff.p,ctx->levelCnt.p,ctx->nLevels,dc.p,dl.p,dh.p);
  ctx->stats.kernel_launches++;
  KCHECK();
  if (count) CK(cudaMemcpyAsync(count,dc.p,sizeof(uint64_t)*n,cudaMemcpyDeviceToHost,st));
  if (lo) CK(cudaMemcpyAsync(lo,dl.p,sizeof(double)*n,cudaMemcpyDeviceToHost,st));
  if (hi) CK(cudaMemcpyAsync(hi,dh.p,sizeof(double)*n,cudaMemcpyDeviceToHost,st));
  CK(cudaStreamSynchronize(st));
  ds.release(); dc.release(); dl.release(); dh.release();
  return WB_OK;
}

extern "C" int wb_query_points(wb_ctx *ctx,const wb_shape *shape,uint64_t cap,uint64_t *nOut,uint32_t *pos,uint32_t *idx,
                               double *x,double *y,double *z)
{
  if (!ctx || !shape || !nOut)
    return WB_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (ctx->phase<PH_BUILT)
    return fail(ctx,WB_ERR_STATE,"not built");
  if (ctx->storeHilbert)
    return fail(ctx,WB_ERR_STATE,"classify-only store (second stage of wb_shard_run): it has no canonical order or leaves");
  int rc=checkShapes(ctx,shape,1);
  if (rc)
    return rc;
  const uint64_t nv=ctx->nValid;
  cudaStream_t st=ctx->st;
  // scr0/scr1 are free between the phases (each phase that uses them fills them first)
  WbShapeDev s;
  memcpy(&s,shape,sizeof(s));
  wb_query_flag_kernel<<<gridFor(nv,256),256,0,st>>>(s,ctx->sx.p,ctx->sy.p,ctx->sz.p,nv,ctx->scr0.p);
  ctx->stats.kernel_launches++;
  CK(wb_exclusive_scan(ctx->scr0.p,ctx->scr1.p,nv,ctx->blockSums.p,ctx->blockSums.cap,st,&ctx->stats.kernel_launches));
  uint32_t lastPos=0,lastFlag=0;
  CK(cudaMemcpyAsync(&lastPos,ctx->scr1.p+nv-1,sizeof(uint32_t),cudaMemcpyDeviceToHost,st));
  CK(cudaMemcpyAsync(&lastFlag,ctx->scr0.p+nv-1,sizeof(uint32_t),cudaMemcpyDeviceToHost,st));
  CK(cudaStreamSynchronize(st));
  const uint64_t total=(uint64_t)lastPos+lastFlag;
  *nOut=total;
  const uint64_t m=std::min(total,cap);
  if (m && (pos || idx || x || y || z))
  {
    DevBuf<uint32_t> dp,di;
    DevBuf<double> dx,dy,dz;
    CK(dp.ensure(m)); CK(di.ensure(m)); CK(dx.ensure(m)); CK(dy.ensure(m)); CK(dz.ensure(m));
    wb_query_emit_kernel<<<gridFor(nv,256),256,0,st>>>(ctx->scr0.p,ctx->scr1.p,nv,ctx->perm,ctx->sx.p,ctx->sy.p,ctx->sz.p,
                                                      m,dp.p,di.p,dx.p,dy.p,dz.p);
    ctx->stats.kernel_launches++;
    KCHECK();
    if (pos) CK(cudaMemcpyAsync(pos,dp.p,sizeof(uint32_t)*m,cudaMemcpyDeviceToHost,st));
    if (idx) CK(cudaMemcpyAsync(idx,di.p,sizeof(uint32_t)*m,cudaMemcpyDeviceToHost,st));
    if (x) CK(cudaMemcpyAsync(x,dx.p,sizeof(double)*m,cudaMemcpyDeviceToHost,st));
    if (y) CK(cudaMemcpyAsync(y,dy.p,sizeof(double)*m,cudaMemcpyDeviceToHost,st));
    if (z) CK(cudaMemcpyAsync(z,dz.p,sizeof(double)*m,cudaMemcpyDeviceToHost,st));
    CK(cudaStreamSynchronize(st));
    dp.release(); di.release(); dx.release(); dy.release(); dz.release();
  }
  return WB_OK;
}

// ============================================================================ output records (ACT_WRITE)

extern "C" int wb_set_return_zero_rule(wb_ctx *ctx,int keep_all)
{
  if (!ctx)
    return WB_ERR_ARG;
  ctx->keepZeroReturns=keep_all!=0;
  return WB_OK;
}

extern "C" int wb_set_window(wb_ctx *ctx,double x_lo,double x_hi)
{
  if (!ctx || !(x_lo<x_hi))
    return WB_ERR_ARG;
  ctx->windowOn=!(x_lo==-INFINITY && x_hi==INFINITY);
  ctx->windowLo=x_lo;
  ctx->windowHi=x_hi;
  return WB_OK;
}

extern "C" int wb_num_loaded(wb_ctx *ctx,uint64_t *n)
{
  if (!ctx || !n)
    return WB_ERR_ARG;
  *n=ctx->n;
  return WB_OK;
}

extern "C" int wb_keep_records(wb_ctx *ctx,int keep)
{
  if (!ctx)
    return WB_ERR_ARG;
  if (ctx->n)
    return fail(ctx,WB_ERR_STATE,"wb_keep_records comes before the first wb_add_las");
  ctx->keepRecords=keep!=0;
  return WB_OK;
}

extern "C" int wb_get_duplicates(wb_ctx *ctx,uint32_t *dup,uint32_t *rep,uint64_t cap)
{
  if (!ctx || (cap && (!dup || !rep)))
    return WB_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (ctx->phase<PH_BUILT)
    return fail(ctx,WB_ERR_STATE,"not built");
  if (cap<ctx->nDup)
    return fail(ctx,WB_ERR_ARG,"duplicate buffers too small");
  if (ctx->nDup)
  {
    CK(cudaMemcpy(dup,ctx->dupIn.p,sizeof(uint32_t)*ctx->nDup,cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(rep,ctx->dupRep.p,sizeof(uint32_t)*ctx->nDup,cudaMemcpyDeviceToHost));
  }
  return WB_OK;
}

namespace
{
int uploadClassLut(wb_ctx *ctx,const uint8_t *classes,int nClasses,int separate)
{
  uint8_t lut[256];
  memset(lut,separate?255:0,sizeof(lut));
  if (separate)
    for (int k=0;k<nClasses;k++)
    {
      if (lut[classes[k]]!=255)
        return fail(ctx,WB_ERR_ARG,"class %d listed twice",(int)classes[k]);
      lut[classes[k]]=(uint8_t)k;
    }
  CK(ctx->encLut.ensure(256));
  CK(cudaMemcpyAsync(ctx->encLut.p,lut,256,cudaMemcpyHostToDevice,ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  return WB_OK;
}
}

extern "C" int wb_leaf_class_counts(wb_ctx *ctx,const uint8_t *classes,int nClasses,int separate,uint32_t *counts)
{
  if (!ctx || !counts || (separate && (!classes || nClasses<1 || nClasses>255)))
    return WB_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (ctx->phase<PH_CLASSIFIED)
    return fail(ctx,WB_ERR_STATE,"not classified");
  if (ctx->storeHilbert)
    return fail(ctx,WB_ERR_STATE,"classify-only store (second stage of wb_shard_run): it has no canonical order or leaves");
  const int K=separate?nClasses:1;
  int rc=uploadClassLut(ctx,classes,nClasses,separate);
  if (rc)
    return rc;
  const uint64_t m=(uint64_t)ctx->nLeaves*K;
  CK(ctx->encCounts.ensure(m+1));
  wb_leaf_class_counts_kernel<<<gridFor((uint64_t)ctx->nLeaves*32,256),256,0,ctx->st>>>(
      ctx->leaves.p,ctx->nLeaves,ctx->labelSorted.p,ctx->encLut.p,K,ctx->encCounts.p);
  ctx->stats.kernel_launches++;
  KCHECK();
  CK(cudaMemcpyAsync(counts,ctx->encCounts.p,sizeof(uint32_t)*m,cudaMemcpyDeviceToHost,ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  return WB_OK;
}

extern "C" int wb_encode(wb_ctx *ctx,const wb_out_spec *spec,const uint64_t *dest,const uint32_t *fileOf,uint32_t nFiles,
                         uint8_t *out,uint64_t outBytes,wb_file_stats *stats)
{
  if (!ctx || !spec || !dest || !fileOf || !nFiles || !stats)
    return WB_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (ctx->phase<PH_CLASSIFIED)
    return fail(ctx,WB_ERR_STATE,"not classified");
  if (ctx->storeHilbert)
    return fail(ctx,WB_ERR_STATE,"classify-only store (second stage of wb_shard_run): it has no canonical order or leaves");
  static const int len[11]={20,28,26,34,57,63,30,36,38,59,67};
  if (spec->format<0 || spec->format>10 || spec->rec_len!=len[spec->format] || spec->rec_len>WB_DEC_MAXLEN)
    return fail(ctx,WB_ERR_FORMAT,"output format %d / record length %d not supported",spec->format,spec->rec_len);
  if (spec->separate && (spec->n_classes<1 || spec->n_classes>255))
    return fail(ctx,WB_ERR_ARG,"1..255 classes");
  if (ctx->recSegs.size()!=ctx->segs.size())
    return fail(ctx,WB_ERR_STATE,"internal: record segments");
  for (size_t i=0;i<ctx->segs.size();i++)
    if (ctx->segs[i].count && !ctx->recSegs[i].recs)
      return fail(ctx,WB_ERR_STATE,"the records were not kept: call wb_keep_records(ctx,1) before wb_add_las");
  if (outBytes&1)
    return fail(ctx,WB_ERR_ARG,"odd output size");
  cudaStream_t st=ctx->st;
  const int K=spec->separate?spec->n_classes:1;
  int rc=uploadClassLut(ctx,spec->classes,spec->n_classes,spec->separate);
  if (rc)
    return rc;
  const uint64_t m=(uint64_t)ctx->nLeaves*K,nv=ctx->nValid;
  CK(cudaEventRecord(ctx->evA,st));
  CK(ctx->encDest.ensure(m+1)); CK(ctx->encFile.ensure(m+1));
  CK(ctx->encMinMax.ensure((uint64_t)nFiles*6)); CK(ctx->encCount.ensure((uint64_t)nFiles*16));
  CK(ctx->outArena.ensure(outBytes+64));
  CK(cudaMemcpyAsync(ctx->encDest.p,dest,sizeof(uint64_t)*m,cudaMemcpyHostToDevice,st));
  CK(cudaMemcpyAsync(ctx->encFile.p,fileOf,sizeof(uint32_t)*m,cudaMemcpyHostToDevice,st));
  {
    std::vector<int> mm((size_t)nFiles*6);
    for (uint32_t f=0;f<nFiles;f++)
      for (int k=0;k<6;k++)
        mm[(size_t)f*6+k]=k<3?0x7fffffff:(int)0x80000000;
    CK(cudaMemcpyAsync(ctx->encMinMax.p,mm.data(),sizeof(int)*mm.size(),cudaMemcpyHostToDevice,st));
    CK(cudaMemsetAsync(ctx->encCount.p,0,sizeof(unsigned long long)*nFiles*16,st));
    CK(cudaStreamSynchronize(st));
  }
  {
    WbRecSegs hr;
    memset(&hr,0,sizeof(hr));
    for (size_t i=0;i<ctx->recSegs.size();i++)
      hr.s[i]=ctx->recSegs[i];
    CK(ctx->drsegs.ensure(1));
    CK(cudaMemcpyAsync(ctx->drsegs.p,&hr,sizeof(WbRecSeg)*std::max<size_t>(1,ctx->recSegs.size()),cudaMemcpyHostToDevice,st));
    CK(cudaStreamSynchronize(st));
  }
  // whose attributes a stored point carries: its own record, or the last record at the same XYZ
  const uint32_t *src=ctx->perm;
  if (ctx->nDup)
  {
    CK(ctx->invPerm.ensure(ctx->n)); CK(ctx->attrSrc.ensure(nv));
    wb_attr_source_kernel<<<gridFor(nv,256),256,0,st>>>(ctx->perm,nv,ctx->invPerm.p,ctx->attrSrc.p);
    wb_attr_last_kernel<<<gridFor(ctx->nDup,256),256,0,st>>>(ctx->dupIn.p,ctx->dupRep.p,ctx->nDup,ctx->invPerm.p,ctx->attrSrc.p);
    ctx->stats.kernel_launches+=2;
    src=ctx->attrSrc.p;
  }
  WbOutSpec ds;
  ds.fmt=spec->format; ds.recLen=spec->rec_len; ds.nClasses=spec->n_classes; ds.separate=spec->separate;
  for (int k=0;k<3;k++)
  {
    ds.scale[k]=spec->scale[k];
    ds.offset[k]=spec->offset[k];
  }
  ds.unit=spec->unit;
  size_t smem=((nFiles<=WB_ENC_SMEM_FILES?(size_t)nFiles*WB_ENC_ACC:0)+(size_t)WB_ENC_WARPS*K)*sizeof(int);
  wb_encode_kernel<<<(unsigned)wb_div_up(ctx->nLeaves,WB_ENC_WARPS),WB_ENC_WARPS*32,smem,st>>>(
      ctx->leaves.p,ctx->nLeaves,ctx->sx.p,ctx->sy.p,ctx->sz.p,ctx->labelSorted.p,src,ctx->dsegs.p,ctx->drsegs.p,
      ds,ctx->encLut.p,ctx->encDest.p,ctx->encFile.p,nFiles,ctx->outArena.p,ctx->encMinMax.p,ctx->encCount.p);
  ctx->stats.kernel_launches++;
  KCHECK();
  CK(cudaEventRecord(ctx->evB,st));
  if (out)
    CK(cudaMemcpyAsync(out,ctx->outArena.p,outBytes,cudaMemcpyDeviceToHost,st));
  ctx->outBytes=outBytes;
  std::vector<int> mm((size_t)nFiles*6);
  std::vector<unsigned long long> cnt((size_t)nFiles*16);
  CK(cudaMemcpyAsync(mm.data(),ctx->encMinMax.p,sizeof(int)*mm.size(),cudaMemcpyDeviceToHost,st));
  CK(cudaMemcpyAsync(cnt.data(),ctx->encCount.p,sizeof(unsigned long long)*cnt.size(),cudaMemcpyDeviceToHost,st));
  CK(cudaEventRecord(ctx->evC,st));
  CK(cudaStreamSynchronize(st));
  for (uint32_t f=0;f<nFiles;f++)
  {
    for (int r=0;r<16;r++)
      stats[f].n_points[r]=cnt[(size_t)f*16+r];
    for (int k=0;k<3;k++)
    {
      stats[f].imin[k]=mm[(size_t)f*6+k];
      stats[f].imax[k]=mm[(size_t)f*6+3+k];
    }
    stats[f].pad_[0]=stats[f].pad_[1]=0;
  }
  ctx->stats.ms_encode=elapsed(ctx->evA,ctx->evB);
  ctx->stats.ms_encode_d2h=elapsed(ctx->evB,ctx->evC);
  return WB_OK;
}

extern "C" int wb_census(wb_ctx *ctx,wb_census_result *out,uint64_t *missing,uint64_t cap)
{
  if (!ctx || !out || (cap && !missing))
    return WB_ERR_ARG;
  cudaSetDevice(ctx->device);
  memset(out,0,sizeof(*out));
  if (ctx->phase<PH_BUILT)
    return fail(ctx,WB_ERR_STATE,"not built");
  if (ctx->recSegs.size()!=ctx->segs.size())
    return fail(ctx,WB_ERR_STATE,"internal: record segments");
  for (size_t i=0;i<ctx->segs.size();i++)
    if (ctx->segs[i].count && !ctx->recSegs[i].recs)
      return fail(ctx,WB_ERR_STATE,"the records were not kept: call wb_keep_records(ctx,1) before wb_add_las");
  cudaStream_t st=ctx->st;
  const uint64_t nv=ctx->nValid;
  out->n_stored=nv;
  {
    WbRecSegs hr;
    memset(&hr,0,sizeof(hr));
    for (size_t i=0;i<ctx->recSegs.size();i++)
      hr.s[i]=ctx->recSegs[i];
    CK(ctx->drsegs.ensure(1));
    CK(cudaMemcpyAsync(ctx->drsegs.p,&hr,sizeof(WbRecSeg)*std::max<size_t>(1,ctx->recSegs.size()),cudaMemcpyHostToDevice,st));
    CK(cudaStreamSynchronize(st));
  }
  const uint32_t *src=ctx->perm;
  if (ctx->nDup)
  {
    // the store holds the LAST record put at a location (octree.cpp:620-662)
    CK(ctx->invPerm.ensure(ctx->n)); CK(ctx->attrSrc.ensure(nv));
    wb_attr_source_kernel<<<gridFor(nv,256),256,0,st>>>(ctx->perm,nv,ctx->invPerm.p,ctx->attrSrc.p);
    wb_attr_last_kernel<<<gridFor(ctx->nDup,256),256,0,st>>>(ctx->dupIn.p,ctx->dupRep.p,ctx->nDup,ctx->invPerm.p,ctx->attrSrc.p);
    ctx->stats.kernel_launches+=2;
    src=ctx->attrSrc.p;
  }
  TmpBuf<unsigned long long> bits,list;
  TmpBuf<WbCensus> cen;
  CK(cen.ensure(1)); CK(list.ensure(cap+1));
  CK(cudaMemsetAsync(cen.p,0,sizeof(WbCensus),st));
  CK(cudaMemsetAsync(ctx->counters.p+15,0,sizeof(unsigned long long),st));
  WbCensus h;
  memset(&h,0,sizeof(h));
  if (nv)
  {
    wb_census_mark_kernel<<<gridFor(nv,256),256,0,st>>>(src,nv,ctx->dsegs.p,ctx->drsegs.p,nullptr,cen.p);
    ctx->stats.kernel_launches++;
    KCHECK();
    CK(cudaMemcpyAsync(&h,cen.p,sizeof(h),cudaMemcpyDeviceToHost,st));
    CK(cudaStreamSynchronize(st));
  }
  if (h.notInteger)
  {
    out->status=-1;
    return WB_OK;
  }
  unsigned long long nMissing=0;
  if (h.maxPlusOne)
  {
    const uint64_t nWords=wb_div_up(h.maxPlusOne,64);
    CK(bits.ensure(nWords));
    CK(cudaMemsetAsync(bits.p,0,nWords*sizeof(unsigned long long),st));
    wb_census_mark_kernel<<<gridFor(nv,256),256,0,st>>>(src,nv,ctx->dsegs.p,ctx->drsegs.p,bits.p,cen.p);
    wb_census_missing_kernel<<<gridFor(nWords,256),256,0,st>>>(bits.p,h.maxPlusOne,ctx->counters.p+15,list.p,cap);
    ctx->stats.kernel_launches+=2;
    KCHECK();
    CK(cudaMemcpyAsync(&h,cen.p,sizeof(h),cudaMemcpyDeviceToHost,st));
    CK(cudaMemcpyAsync(&nMissing,ctx->counters.p+15,sizeof(nMissing),cudaMemcpyDeviceToHost,st));
    CK(cudaStreamSynchronize(st));
  }
  out->status=h.duplicate?1:0;
  out->max_point=h.maxPlusOne;
  out->n_duplicate=h.duplicate;
  out->n_missing=nMissing;
  const uint64_t k=std::min<uint64_t>(cap,nMissing);
  if (k)
  {
    CK(cudaMemcpy(missing,list.p,k*sizeof(uint64_t),cudaMemcpyDeviceToHost));
    std::sort(missing,missing+k);
  }
  return WB_OK;
}

extern "C" int wb_write_encoded(wb_ctx *ctx,int fd,uint64_t filePos,uint64_t arenaOff,uint64_t bytes)
// Stream a span of the records wb_encode left in device memory into an open file: D2H into the
// pinned ring, pwrite on worker threads (the page-cache copy is the slow part and runs 8 wide).
{
  if (!ctx || fd<0)
    return WB_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (arenaOff+bytes>ctx->outBytes)
    return fail(ctx,WB_ERR_ARG,"span outside the encoded records");
  if (!bytes)
    return WB_OK;
  const int T=WB_READ_THREADS;
  const uint64_t chunk=std::max<uint64_t>(ctx->readBufBytes>=((uint64_t)8<<20)?ctx->readBufBytes:0,(uint64_t)16<<20);
  int rc=ensureRing(ctx,chunk);
  if (rc)
    return rc;
  const uint64_t nChunks=wb_div_up(bytes,chunk);
  // Default: pwrite (errors come back as codes).  WB_WRITE_MODE=mmap stores through a shared mapping
  // instead, which avoids the inode lock that serialises buffered writes to one file; on the test
  // box (ext4 on virtio) both reach the same 4.7 GB/s, bound by page-cache allocation.
  uint8_t *mapped=nullptr;
  uint64_t mapLen=0,mapSkew=0;
  {
    const char *mode=getenv("WB_WRITE_MODE");
    struct stat sb;
    if (mode && !strcmp(mode,"mmap") && fstat(fd,&sb)==0 && S_ISREG(sb.st_mode))
    {
      if ((uint64_t)sb.st_size>=filePos+bytes || ftruncate(fd,(off_t)(filePos+bytes))==0)
      {
        const uint64_t page=(uint64_t)sysconf(_SC_PAGESIZE);
        mapSkew=filePos%page;
        mapLen=bytes+mapSkew;
        void *m=mmap(nullptr,mapLen,PROT_READ|PROT_WRITE,MAP_SHARED,fd,(off_t)(filePos-mapSkew));
        if (m!=MAP_FAILED)
          mapped=(uint8_t *)m;
      }
    }
  }
  ReadRing ring;                                       // state: 0 free, 1 filled, -1 write error
  auto worker=[&](int t)
  {
    for (uint64_t k=t;k<nChunks;k+=T)
    {
      {
        std::unique_lock<std::mutex> lk(ring.m);
        ring.cv.wait(lk,[&]{ return ring.state[t]==1 || ring.stop; });
        if (ring.state[t]!=1)
          return;
      }
      uint64_t want=std::min(chunk,bytes-k*chunk),done=0;
      if (mapped)
      {
        memcpy(mapped+mapSkew+k*chunk,ctx->readBuf[t],want);
        done=want;
      }
      while (done<want)
      {
        ssize_t w=pwrite(fd,ctx->readBuf[t]+done,want-done,(off_t)(filePos+k*chunk+done));
        if (w<=0)
          break;
        done+=(uint64_t)w;
      }
      {
        std::lock_guard<std::mutex> lk(ring.m);
        ring.state[t]=done==want?0:-1;
      }
      ring.cv.notify_all();
    }
  };
  std::vector<std::thread> threads;
  for (int t=0;t<T && (uint64_t)t<nChunks;t++)
    threads.emplace_back(worker,t);
  cudaError_t ce=cudaSuccess;
  bool ioError=false;
  for (uint64_t k=0;k<nChunks && ce==cudaSuccess && !ioError;k++)
  {
    const int t=(int)(k%T);
    {
      std::unique_lock<std::mutex> lk(ring.m);
      ring.cv.wait(lk,[&]{ return ring.state[t]!=1; });
      if (ring.state[t]<0)
      {
        ioError=true;
        break;
      }
    }
    ce=cudaMemcpyAsync(ctx->readBuf[t],ctx->outArena.p+arenaOff+k*chunk,std::min(chunk,bytes-k*chunk),
                       cudaMemcpyDeviceToHost,ctx->stCopy);
    if (ce==cudaSuccess)
      ce=cudaStreamSynchronize(ctx->stCopy);
    if (ce!=cudaSuccess)
      break;
    {
      std::lock_guard<std::mutex> lk(ring.m);
      ring.state[t]=1;
    }
    ring.cv.notify_all();
  }
  {
    std::unique_lock<std::mutex> lk(ring.m);
    ring.cv.wait(lk,[&]
    {
      for (int t=0;t<T;t++)
        if (ring.state[t]==1)
          return false;
      return true;
    });
    for (int t=0;t<T;t++)
      if (ring.state[t]<0)
        ioError=true;
    ring.stop=true;
  }
  ring.cv.notify_all();
  for (auto &th:threads)
    th.join();
  if (mapped)
    munmap(mapped,mapLen);
  if (ce!=cudaSuccess)
    return fail(ctx,WB_ERR_CUDA,"%s",cudaGetErrorString(ce));
  if (ioError)
    return fail(ctx,WB_ERR_ARG,"write failed (disk full?)");
  return WB_OK;
}

extern "C" int wb_patch_records(wb_ctx *ctx,uint8_t *recs,uint64_t first,uint64_t n,int fmt,int recLen)
{
  if (!ctx || !recs)
    return WB_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (ctx->phase<PH_CLASSIFIED)
    return fail(ctx,WB_ERR_STATE,"not classified");
  if (first+n>ctx->n)
    return fail(ctx,WB_ERR_ARG,"range outside the cloud");
  std::vector<uint8_t> lab(n);
  CK(cudaMemcpy(lab.data(),ctx->labelIn.p+first,n,cudaMemcpyDeviceToHost));
  for (uint64_t i=0;i<n;i++)
  {
    uint8_t *r=recs+i*(uint64_t)recLen;
    if (fmt<6)
      r[15]=(uint8_t)((r[15]&0xe0)|(lab[i]&31));       // writePoint, las.cpp:848
    else
      r[16]=lab[i];                                      // las.cpp:857
  }
  return WB_OK;
}

extern "C" int wb_test_math(wb_ctx *ctx,uint64_t n,const double *y,const double *x,int32_t *ao,double *ho,int32_t *so)
{
  if (!ctx || !y || !x || !ao || !ho || !so)
    return WB_ERR_ARG;
  cudaSetDevice(ctx->device);
  DevBuf<double> dy,dx,dh;
  DevBuf<int> da,ds;
  CK(dy.ensure(n)); CK(dx.ensure(n)); CK(dh.ensure(n)); CK(da.ensure(n)); CK(ds.ensure(n));
  CK(cudaMemcpy(dy.p,y,n*sizeof(double),cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dx.p,x,n*sizeof(double),cudaMemcpyHostToDevice));
  wb_test_math_kernel<<<gridFor(n,256),256,0,ctx->st>>>(dy.p,dx.p,n,da.p,dh.p,ds.p);
  ctx->stats.kernel_launches++;
  KCHECK();
  CK(cudaStreamSynchronize(ctx->st));
  CK(cudaMemcpy(ao,da.p,n*sizeof(int),cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ho,dh.p,n*sizeof(double),cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(so,ds.p,n*sizeof(int),cudaMemcpyDeviceToHost));
  dy.release(); dx.release(); dh.release(); da.release(); ds.release();
  return WB_OK;
}

extern "C" int wb_run(wb_ctx *ctx)
{
  int rc;
  if ((rc=wb_build(ctx))) return rc;
  if ((rc=wb_scan(ctx))) return rc;
  if ((rc=wb_postscan(ctx))) return rc;
  return wb_classify(ctx);
}

extern "C" int wb_get_stats(wb_ctx *ctx,wb_stats *out)
{
  if (!ctx || !out)
    return WB_ERR_ARG;
  *out=ctx->stats;
  return WB_OK;
}

extern "C" int wb_sync(wb_ctx *ctx)
{
  if (!ctx)
    return WB_ERR_ARG;
  cudaSetDevice(ctx->device);
  CK(cudaStreamSynchronize(ctx->st));
  CK(cudaStreamSynchronize(ctx->stCopy));
  return WB_OK;
}

extern "C" int wb_mark(wb_ctx *ctx,int slot)
// A timing event on the compute stream, behind everything issued so far (both streams: the copy
// stream is joined first, so a mark after wb_add_las also covers its transfers).
{
  if (!ctx || slot<0 || slot>=WB_MARKS)
    return WB_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (!ctx->evMark[slot])
    CK(cudaEventCreate(&ctx->evMark[slot]));
  if (!ctx->evJoin)
    CK(cudaEventCreateWithFlags(&ctx->evJoin,cudaEventDisableTiming));
  CK(cudaEventRecord(ctx->evJoin,ctx->stCopy));
  CK(cudaStreamWaitEvent(ctx->st,ctx->evJoin,0));
  CK(cudaEventRecord(ctx->evMark[slot],ctx->st));
  return WB_OK;
}

extern "C" int wb_mark_elapsed(wb_ctx *ctx,int from,int to,double *ms)
{
  if (!ctx || !ms || from<0 || from>=WB_MARKS || to<0 || to>=WB_MARKS)
    return WB_ERR_ARG;
  if (!ctx->evMark[from] || !ctx->evMark[to])
    return fail(ctx,WB_ERR_STATE,"wb_mark_elapsed: mark %d or %d was never recorded",from,to);
  cudaSetDevice(ctx->device);
  CK(cudaEventSynchronize(ctx->evMark[to]));
  float f=0;
  CK(cudaEventElapsedTime(&f,ctx->evMark[from],ctx->evMark[to]));
  *ms=f;
  return WB_OK;
}

// ============================================================================ host helpers

extern "C" int wb_host_alloc(void **p,uint64_t bytes)
{
  if (!p)
    return WB_ERR_ARG;
  return cudaHostAlloc(p,(size_t)bytes,cudaHostAllocDefault)==cudaSuccess?WB_OK:WB_ERR_NOMEM;
}

extern "C" int wb_host_free(void *p)
{
  return cudaFreeHost(p)==cudaSuccess?WB_OK:WB_ERR_CUDA;
}

extern "C" int wb_size_fit(const double *corners,int n,double center[3],double *side)
{
  if (!corners || !center || !side)
    return WB_ERR_ARG;
  wbhost::sizeFit(corners,n,center,side);
  return WB_OK;
}

extern "C" int wb_bbox_cube(const double *corners,int n,double cube[4])
{
  if (!corners || !cube)
    return WB_ERR_ARG;
  wbhost::bboxCube(corners,n,cube);
  return WB_OK;
}

extern "C" int wb_bound_rect(const double *corners,int n,double box[6])
{
  if (!corners || !box)
    return WB_ERR_ARG;
  wbhost::boundRect(corners,n,box);
  return WB_OK;
}

extern "C" int wb_snake_set_size(double cubeSide,double tileSize,double *spacing,int *lo,int *hi)
{
  if (!spacing || !lo || !hi)
    return WB_ERR_ARG;
  return wbhost::snakeSetSize(cubeSide,tileSize,spacing,lo,hi);
}

extern "C" int wb_ldecimal(double x,char *buf,int buflen)
{
  std::string s=wbhost::ldecimal(x);
  if (!buf || (int)s.size()+1>buflen)
    return WB_ERR_ARG;
  memcpy(buf,s.c_str(),s.size()+1);
  return (int)s.size();
}

extern "C" int wb_format_dump(const wb_leaf *leaves,uint64_t n,char *buf,uint64_t buflen)
// OctStore::dump / OctBuffer::dump text (octree.cpp:673-689, 888-891); returns the length
{
  if ((!leaves && n) || !buf)
    return WB_ERR_ARG;
  // shortest-round-trip decimals cost a few microseconds per line: format ranges of leaves in parallel
  unsigned hw=std::thread::hardware_concurrency();
  const uint64_t nt=std::max<uint64_t>(1,std::min<uint64_t>(std::min<unsigned>(hw?hw:1,16),n/4096));
  std::vector<std::string> part(nt);
  std::vector<unsigned long long> sub(nt,0);
  auto work=[&](uint64_t t)
  {
    std::string &o=part[t];
    const uint64_t a=n*t/nt,b=n*(t+1)/nt;
    o.reserve((b-a)*64+64);
    for (uint64_t i=a;i<b;i++)
    {
      o+="("+wbhost::ldecimal(leaves[i].cx)+","+wbhost::ldecimal(leaves[i].cy)+","+wbhost::ldecimal(leaves[i].cz)+")\xc2\xb1";
      o+=wbhost::ldecimal(leaves[i].half)+" "+std::to_string(leaves[i].count)+" points\n";
      sub[t]+=leaves[i].count;
    }
  };
  {
    std::vector<std::thread> th;
    for (uint64_t t=1;t<nt;t++)
      th.emplace_back(work,t);
    work(0);
    for (auto &x:th)
      x.join();
  }
  std::string out;
  unsigned long long total=0;
  {
    size_t len=64;
    for (auto &o:part)
      len+=o.size();
    out.reserve(len);
  }
  for (uint64_t t=0;t<nt;t++)
  {
    out+=part[t];
    total+=sub[t];
  }
  out+=std::to_string(total)+" total points\n";
  if (out.size()+1>buflen)
    return WB_ERR_ARG;
  memcpy(buf,out.c_str(),out.size()+1);
  return (int)out.size();
}

#include "wb_shard.cuh"
