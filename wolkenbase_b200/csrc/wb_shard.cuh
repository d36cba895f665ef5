// wb_shard.cuh — one cloud over several GPUs: x-strips, halo exchange, one pipeline per rank.
// Included at the end of wolken_b200.cu (it works on wb_ctx's internals).
//
// The reference has no distributed mode (one std::thread pool, threads.cpp:91-113); the analogue of
// startThreads(n) here is n ranks, one GPU each, that run the SAME phases on their own strip of the cloud and
// exchange only what crosses a strip border:
//
//   geometry   every rank contributes the bounding box of its files' header corners; Octree::sizeFit and the
//              flowsnake cube depend only on the overall extremes (wb_host.h), so Morton keys, tile numbers and
//              tile centres are identical everywhere and equal to the single-GPU ones;
//   halo 1     points within 2.5 tile spacings of another rank's strip go there: every tile whose cylinder
//              (radius 41/71 spacing) holds one of a rank's own points is then complete on that rank, so the
//              tile scan needs no exchange of results at all;
//   tile grid  postscanCylinder's ray walks (scan.cpp:142-179) only read "populated / populated tree tile" of other
//              tiles: one byte per lattice cell, each rank filling the cells of the tiles whose centre lies in its
//              ownership interval, combined with one all-reduce (max) — 1 B per cell instead of the 16 B per tile
//              x 3 all-reduces of round 1;
//   halo 2     a point Q can lie in the hyperboloid of P only if dist_xy <= sqrt(dz^2+2 por dz)/slope with
//              dz <= zmax-thickness-Q.z and por <= max hyperboloidSize x slope^2 (shape.cpp:119-135): Q goes to every
//              rank whose strip is that close; the rank rebuilds over own+halo and classifies its own points;
//   order      local input order is [halo from lower ranks | own | halo from higher ranks], each sender's points in
//              their own order: the canonical order (Morton key, then input index) restricted to a rank equals
//              the global one, which the tile scan's pairwise sums and bottom2 depend on.
//
// Halo rows carry the sender's integers and the index of the sender's segment; every rank knows every other
// rank's segments (scale, offset, unit per input file), so coordinates are rebuilt with the SENDER's header,
// exactly as the owner computes them.  Records dropped by the return-number rule are never sent.
//
// Transport: wb_comm has three bodies.  NCCL (ncclSend/ncclRecv groups, ncclAllGather, ncclAllReduce, resolved
// with dlopen so that the library loads on a box without libnccl) is the product path, for ranks in different
// processes (torchrun) or threads (wolkencli --gpus N).  LOCAL exchanges between the threads of one process with
// plain copies and a barrier; CUSTOM calls back into the host program.  The last two exist so that the whole
// sharded pipeline can run under the CPU emulator of tests/simt (gloo stands in for NCCL there).
#pragma once
#include <dlfcn.h>
#include <chrono>

// ============================================================================ kernels

#define WB_HALO_THREADS 512
#define WB_HALO_MAXDEST 4

struct WbHaloArgs
{
  int nDest;
  int mode;                       // 1: fixed radius, 2: radius from the point's height
  double lo[WB_HALO_MAXDEST],hi[WB_HALO_MAXDEST];
  double radius;                  // mode 1
  double zTop,porMax,slope;       // mode 2: zmax-thickness, largest polar radius, maxSlope
};

__device__ __forceinline__ uint32_t wb_halo_flags(const WbHaloArgs &a,double x,double z)
{
  double r=a.radius;
  if (a.mode==2)
  {
    double dz=fmax(a.zTop-z,0.0);
    r=sqrt(dz*dz+2.0*a.porMax*dz)/a.slope*(1+1e-9)+1e-6;
  }
  uint32_t f=0;
  for (int d=0;d<a.nDest;d++)
    if (x>=a.lo[d]-r && x<=a.hi[d]+r)
      f|=1u<<d;
  return f;
}

__global__ void __launch_bounds__(WB_HALO_THREADS)
wb_halo_count_kernel(const int *__restrict__ xi,const int *__restrict__ zi,const uint8_t *__restrict__ ret,
                     unsigned long long first,unsigned long long cnt,WbSegment seg,WbHaloArgs a,
                     uint32_t *__restrict__ blockCounts,uint32_t blockBase,uint32_t totalBlocks)
// blockCounts[d*totalBlocks+blockBase+block] = points of this block that go to destination d
{
  __shared__ uint32_t sc[WB_HALO_MAXDEST];
  if (threadIdx.x<WB_HALO_MAXDEST)
    sc[threadIdx.x]=0;
  __syncthreads();
  unsigned long long t=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  uint32_t f=0;
  if (t<cnt && ret[first+t])
    f=wb_halo_flags(a,wb_coord(seg.offset[0],seg.scale[0],xi[first+t],seg.unit),
                    wb_coord(seg.offset[2],seg.scale[2],zi[first+t],seg.unit));
  for (int d=0;d<a.nDest;d++)
  {
    uint32_t m=__ballot_sync(WB_FULL,(f>>d)&1);
    if ((threadIdx.x&31)==0 && m)
      atomicAdd(&sc[d],(uint32_t)__popc(m));
  }
  __syncthreads();
  if (threadIdx.x<(unsigned)a.nDest)
    blockCounts[(unsigned long long)threadIdx.x*totalBlocks+blockBase+blockIdx.x]=sc[threadIdx.x];
}

__global__ void __launch_bounds__(WB_HALO_THREADS)
wb_halo_scatter_kernel(const int *__restrict__ xi,const int *__restrict__ yi,const int *__restrict__ zi,
                       const uint8_t *__restrict__ cls,const uint8_t *__restrict__ ret,
                       unsigned long long first,unsigned long long cnt,WbSegment seg,int segIndex,WbHaloArgs a,
                       const uint32_t *__restrict__ blockOff,uint32_t blockBase,uint32_t totalBlocks,
                       int4 *__restrict__ rows)
// rows of destination d start at blockOff[d*totalBlocks] (one exclusive scan over the whole count table): dest-major,
// then segment, then input order — the order the receiver keeps
{
  __shared__ uint32_t warpCnt[WB_HALO_MAXDEST][WB_HALO_THREADS/32];
  unsigned long long t=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  const int lane=threadIdx.x&31,wi=threadIdx.x>>5;
  uint32_t f=0;
  int x=0,y=0,z=0;
  uint32_t w3=0;
  if (t<cnt && ret[first+t])
  {
    x=xi[first+t]; y=yi[first+t]; z=zi[first+t];
    w3=(uint32_t)cls[first+t]|((uint32_t)ret[first+t]<<8)|((uint32_t)segIndex<<16);
    f=wb_halo_flags(a,wb_coord(seg.offset[0],seg.scale[0],x,seg.unit),wb_coord(seg.offset[2],seg.scale[2],z,seg.unit));
  }
  uint32_t before[WB_HALO_MAXDEST];
  for (int d=0;d<a.nDest;d++)
  {
    uint32_t m=__ballot_sync(WB_FULL,(f>>d)&1);
    before[d]=__popc(m&((1u<<lane)-1));
    if (lane==0)
      warpCnt[d][wi]=__popc(m);
  }
  __syncthreads();
  for (int d=0;d<a.nDest;d++)
    if ((f>>d)&1)
    {
      uint32_t o=blockOff[(unsigned long long)d*totalBlocks+blockBase+blockIdx.x]+before[d];
      for (int w=0;w<wi;w++)
        o+=warpCnt[d][w];
      rows[o]=make_int4(x,y,z,(int)w3);
    }
}

__global__ void __launch_bounds__(256)
wb_halo_unpack_kernel(const int4 *__restrict__ rows,unsigned long long n,
                      int *__restrict__ ox,int *__restrict__ oy,int *__restrict__ oz,uint8_t *__restrict__ oc,
                      uint8_t *__restrict__ oret)
{
  unsigned long long i=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  if (i>=n)
    return;
  int4 r=rows[i];
  ox[i]=r.x; oy[i]=r.y; oz[i]=r.z;
  oc[i]=(uint8_t)(r.w&255);
  oret[i]=(uint8_t)((r.w>>8)&255);
}

__global__ void __launch_bounds__(256)
wb_copy_columns_kernel(const int *__restrict__ x,const int *__restrict__ y,const int *__restrict__ z,
                       const uint8_t *__restrict__ c,const uint8_t *__restrict__ r,unsigned long long n,
                       int *__restrict__ ox,int *__restrict__ oy,int *__restrict__ oz,uint8_t *__restrict__ oc,
                       uint8_t *__restrict__ oret)
{
  unsigned long long i=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  if (i>=n)
    return;
  ox[i]=x[i]; oy[i]=y[i]; oz[i]=z[i];
  oc[i]=c[i];
  oret[i]=r[i];
}

__global__ void __launch_bounds__(256)
wb_max_u8_kernel(uint8_t *__restrict__ a,const uint8_t *__restrict__ b,unsigned long long n)
{
  unsigned long long i=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  if (i<n && b[i]>a[i])
    a[i]=b[i];
}

__global__ void __launch_bounds__(256)
wb_sorted_labels_kernel(const uint8_t *__restrict__ labelIn,const uint32_t *__restrict__ perm,unsigned long long n,
                        uint8_t *__restrict__ labelSorted)
// wb_set_labels: the canonical-order copy of labels given in input order
{
  unsigned long long j=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  if (j<n)
    labelSorted[j]=labelIn[perm[j]];
}

// ============================================================================ transport

extern "C"
{
typedef struct wb_nccl_id { char internal[128]; } wb_nccl_id;     // ncclUniqueId
}

namespace
{

struct NcclApi
{
  void *lib=nullptr;
  int (*GetUniqueId)(wb_nccl_id *)=nullptr;
  int (*CommInitRank)(void **,int,wb_nccl_id,int)=nullptr;
  int (*CommDestroy)(void *)=nullptr;
  int (*AllGather)(const void *,void *,size_t,int,void *,cudaStream_t)=nullptr;
  int (*AllReduce)(const void *,void *,size_t,int,int,void *,cudaStream_t)=nullptr;
  int (*Send)(const void *,size_t,int,int,void *,cudaStream_t)=nullptr;
  int (*Recv)(void *,size_t,int,int,void *,cudaStream_t)=nullptr;
  int (*GroupStart)()=nullptr;
  int (*GroupEnd)()=nullptr;
  const char *(*GetErrorString)(int)=nullptr;
  std::string err;
};
enum { WB_NCCL_UINT8=1,WB_NCCL_MAX=2 };               // ncclUint8, ncclMax (nccl.h)

NcclApi *ncclApi()
// libnccl.so.2 is taken from wherever the process already has it (a Python process that imported torch has the
// bundled one mapped; dlopen then returns that copy), else from the loader path, else from WB_NCCL_LIB
{
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once,[]
  {
    const char *names[3]={getenv("WB_NCCL_LIB"),"libnccl.so.2","libnccl.so"};
    for (const char *n:names)
      if (n && *n && (api.lib=dlopen(n,RTLD_NOW|RTLD_GLOBAL)))
        break;
    if (!api.lib)
    {
      api.err="libnccl.so.2 not found (set WB_NCCL_LIB)";
      return;
    }
    bool ok=true;
    auto sym=[&](const char *s){ void *p=dlsym(api.lib,s); if (!p) { ok=false; api.err=std::string("missing symbol ")+s; } return p; };
    *(void **)&api.GetUniqueId=sym("ncclGetUniqueId");
    *(void **)&api.CommInitRank=sym("ncclCommInitRank");
    *(void **)&api.CommDestroy=sym("ncclCommDestroy");
    *(void **)&api.AllGather=sym("ncclAllGather");
    *(void **)&api.AllReduce=sym("ncclAllReduce");
    *(void **)&api.Send=sym("ncclSend");
    *(void **)&api.Recv=sym("ncclRecv");
    *(void **)&api.GroupStart=sym("ncclGroupStart");
    *(void **)&api.GroupEnd=sym("ncclGroupEnd");
    *(void **)&api.GetErrorString=sym("ncclGetErrorString");
    if (!ok)
    {
      dlclose(api.lib);
      api.lib=nullptr;
    }
  });
  return &api;
}

} // namespace

struct wb_local_group
// the ranks of one process (one host thread each): a generation barrier and a notice board of pointers
{
  int world=0;
  std::mutex m;
  std::condition_variable cv;
  int arrived=0;
  unsigned long long gen=0;
  std::vector<const void *> ptr;
  std::vector<const uint64_t *> off,cnt;
  void barrier()
  {
    std::unique_lock<std::mutex> lk(m);
    unsigned long long g=gen;
    if (++arrived==world)
    {
      arrived=0;
      gen++;
      cv.notify_all();
    }
    else
      cv.wait(lk,[&]{ return gen!=g; });
  }
};

struct wb_comm
{
  int kind=0;                         // 0 NCCL, 1 LOCAL, 2 CUSTOM
  int rank=0,world=1;
  wb_ctx *ctx=nullptr;
  void *nccl=nullptr;
  wb_local_group *grp=nullptr;
  wb_comm_ops ops{};
  DevBuf<uint8_t> small,snap,tmp;     // staging of host metadata; LOCAL all-reduce
  uint64_t bytesSent=0,bytesReceived=0;
};

namespace
{

#define NCK(call) do { int r_=(call); if (r_!=0) return fail(ctx,WB_ERR_CUDA,"%s: %s",#call,ncclApi()->GetErrorString?ncclApi()->GetErrorString(r_):"NCCL error"); } while (0)

int commAllGather(wb_comm *cm,const void *dSend,void *dRecv,uint64_t bytes)
// every rank's `bytes` bytes, in rank order, to every rank (device buffers; ordered on the context's stream)
{
  wb_ctx *ctx=cm->ctx;
  cudaStream_t st=ctx->st;
  if (cm->kind==0)
  {
    NCK(ncclApi()->AllGather(dSend,dRecv,(size_t)bytes,WB_NCCL_UINT8,cm->nccl,st));
    CK(cudaStreamSynchronize(st));     // see commAllToAllV
    return WB_OK;
  }
  if (cm->kind==2)
  {
    CK(cudaStreamSynchronize(st));
    if (cm->ops.all_gather(cm->ops.user,dSend,dRecv,bytes))
      return fail(ctx,WB_ERR_CUDA,"all_gather callback failed");
    return WB_OK;
  }
  wb_local_group *g=cm->grp;
  CK(cudaStreamSynchronize(st));
  g->ptr[cm->rank]=dSend;
  g->barrier();
  for (int k=0;k<cm->world;k++)
    CK(cudaMemcpyAsync((uint8_t *)dRecv+(uint64_t)k*bytes,g->ptr[k],(size_t)bytes,cudaMemcpyDefault,st));
  CK(cudaStreamSynchronize(st));
  g->barrier();
  return WB_OK;
}

int commAllToAllV(wb_comm *cm,const void *dSend,const uint64_t *sOff,const uint64_t *sCnt,
                  void *dRecv,const uint64_t *rOff,const uint64_t *rCnt)
// bytes sCnt[k] at dSend+sOff[k] go to rank k, which receives them at dRecv+rOff[me] (rCnt[k] = what k sends me)
{
  wb_ctx *ctx=cm->ctx;
  cudaStream_t st=ctx->st;
  for (int k=0;k<cm->world;k++)
  {
    cm->bytesSent+=sCnt[k];
    cm->bytesReceived+=rCnt[k];
  }
  if (cm->kind==0)
  {
    NcclApi *N=ncclApi();
    NCK(N->GroupStart());
    for (int k=0;k<cm->world;k++)
    {
      if (k==cm->rank)
        continue;
      if (sCnt[k])
        NCK(N->Send((const uint8_t *)dSend+sOff[k],(size_t)sCnt[k],WB_NCCL_UINT8,k,cm->nccl,st));
      if (rCnt[k])
        NCK(N->Recv((uint8_t *)dRecv+rOff[k],(size_t)rCnt[k],WB_NCCL_UINT8,k,cm->nccl,st));
    }
    NCK(N->GroupEnd());
    // Drain before the caller touches the CUDA runtime again: with the ranks as THREADS of one process a later
    // cudaMalloc/cudaFree of this thread (device-synchronising, and serialised with the other threads' calls inside
    // the driver) could otherwise wait for this NCCL kernel while the peer's launch waits for that call — the
    // well-known NCCL + blocking-CUDA-call deadlock.  The stages are separated by a stream sync anyway.
    CK(cudaStreamSynchronize(st));
    return WB_OK;
  }
  if (cm->kind==2)
  {
    CK(cudaStreamSynchronize(st));
    if (cm->ops.all_to_all_v(cm->ops.user,dSend,sOff,sCnt,dRecv,rOff,rCnt))
      return fail(ctx,WB_ERR_CUDA,"all_to_all_v callback failed");
    return WB_OK;
  }
  wb_local_group *g=cm->grp;
  CK(cudaStreamSynchronize(st));
  g->ptr[cm->rank]=dSend;
  g->off[cm->rank]=sOff;
  g->cnt[cm->rank]=sCnt;
  g->barrier();
  for (int k=0;k<cm->world;k++)
    if (k!=cm->rank && rCnt[k])
    {
      if (g->cnt[k][cm->rank]!=rCnt[k])
        return fail(ctx,WB_ERR_STATE,"internal: exchange counts disagree");
      CK(cudaMemcpyAsync((uint8_t *)dRecv+rOff[k],(const uint8_t *)g->ptr[k]+g->off[k][cm->rank],(size_t)rCnt[k],
                         cudaMemcpyDefault,st));
    }
  CK(cudaStreamSynchronize(st));
  g->barrier();
  return WB_OK;
}

int commAllReduceMaxU8(wb_comm *cm,uint8_t *dBuf,uint64_t n)
{
  wb_ctx *ctx=cm->ctx;
  cudaStream_t st=ctx->st;
  if (cm->kind==0)
  {
    NCK(ncclApi()->AllReduce(dBuf,dBuf,(size_t)n,WB_NCCL_UINT8,WB_NCCL_MAX,cm->nccl,st));
    CK(cudaStreamSynchronize(st));     // see commAllToAllV
    return WB_OK;
  }
  if (cm->kind==2)
  {
    CK(cudaStreamSynchronize(st));
    if (cm->ops.all_reduce_max_u8(cm->ops.user,dBuf,n))
      return fail(ctx,WB_ERR_CUDA,"all_reduce_max_u8 callback failed");
    return WB_OK;
  }
  wb_local_group *g=cm->grp;
  CK(cm->snap.ensure(n+1)); CK(cm->tmp.ensure(n+1));
  CK(cudaMemcpyAsync(cm->snap.p,dBuf,(size_t)n,cudaMemcpyDeviceToDevice,st));
  CK(cudaStreamSynchronize(st));
  g->ptr[cm->rank]=cm->snap.p;
  g->barrier();
  for (int k=0;k<cm->world;k++)
    if (k!=cm->rank)
    {
      CK(cudaMemcpyAsync(cm->tmp.p,g->ptr[k],(size_t)n,cudaMemcpyDefault,st));
      wb_max_u8_kernel<<<gridFor(n,256),256,0,st>>>(dBuf,cm->tmp.p,n);
      ctx->stats.kernel_launches++;
    }
  CK(cudaStreamSynchronize(st));
  g->barrier();
  return WB_OK;
}

int commAllGatherHost(wb_comm *cm,const void *send,void *recv,uint64_t bytes)
// small host-side metadata through the device collective
{
  wb_ctx *ctx=cm->ctx;
  CK(cm->small.ensure(bytes*(cm->world+1)+16));
  CK(cudaMemcpyAsync(cm->small.p,send,(size_t)bytes,cudaMemcpyHostToDevice,ctx->st));
  int rc=commAllGather(cm,cm->small.p,cm->small.p+bytes,bytes);
  if (rc)
    return rc;
  CK(cudaMemcpyAsync(recv,cm->small.p+bytes,(size_t)(bytes*cm->world),cudaMemcpyDeviceToHost,ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  return WB_OK;
}

} // namespace

extern "C" int wb_comm_get_id(uint8_t id[WB_COMM_ID_BYTES])
{
  if (!id)
    return WB_ERR_ARG;
  NcclApi *N=ncclApi();
  if (!N->lib)
    return WB_ERR_STATE;
  wb_nccl_id u;
  static_assert(sizeof(u)==WB_COMM_ID_BYTES,"ncclUniqueId is 128 bytes");
  if (N->GetUniqueId(&u)!=0)
    return WB_ERR_CUDA;
  memcpy(id,&u,sizeof(u));
  return WB_OK;
}

extern "C" int wb_comm_init(wb_ctx *ctx,const uint8_t id[WB_COMM_ID_BYTES],int rank,int world,wb_comm **out)
{
  if (!ctx || !id || !out || world<1 || rank<0 || rank>=world)
    return WB_ERR_ARG;
  *out=nullptr;
  if (world>WB_MAX_RANKS)
    return fail(ctx,WB_ERR_ARG,"at most %d ranks",WB_MAX_RANKS);
  NcclApi *N=ncclApi();
  if (!N->lib)
    return fail(ctx,WB_ERR_STATE,"NCCL: %s",N->err.c_str());
  cudaSetDevice(ctx->device);
  wb_nccl_id u;
  memcpy(&u,id,sizeof(u));
  void *c=nullptr;
  NCK(N->CommInitRank(&c,world,u,rank));
  wb_comm *cm=new wb_comm;
  cm->kind=0;
  cm->rank=rank;
  cm->world=world;
  cm->ctx=ctx;
  cm->nccl=c;
  *out=cm;
  return WB_OK;
}

extern "C" int wb_local_group_create(int world,wb_local_group **out)
{
  if (!out || world<1 || world>WB_MAX_RANKS)
    return WB_ERR_ARG;
  wb_local_group *g=new wb_local_group;
  g->world=world;
  g->ptr.assign(world,nullptr);
  g->off.assign(world,nullptr);
  g->cnt.assign(world,nullptr);
  *out=g;
  return WB_OK;
}

extern "C" void wb_local_group_destroy(wb_local_group *g)
{
  delete g;
}

extern "C" int wb_comm_init_local(wb_ctx *ctx,wb_local_group *grp,int rank,wb_comm **out)
{
  if (!ctx || !grp || !out || rank<0 || rank>=grp->world)
    return WB_ERR_ARG;
  wb_comm *cm=new wb_comm;
  cm->kind=1;
  cm->rank=rank;
  cm->world=grp->world;
  cm->ctx=ctx;
  cm->grp=grp;
  *out=cm;
  return WB_OK;
}

extern "C" int wb_comm_init_custom(wb_ctx *ctx,const wb_comm_ops *ops,int rank,int world,wb_comm **out)
{
  if (!ctx || !ops || !out || world<1 || world>WB_MAX_RANKS || rank<0 || rank>=world || !ops->all_gather ||
      !ops->all_to_all_v || !ops->all_reduce_max_u8)
    return WB_ERR_ARG;
  wb_comm *cm=new wb_comm;
  cm->kind=2;
  cm->rank=rank;
  cm->world=world;
  cm->ctx=ctx;
  cm->ops=*ops;
  *out=cm;
  return WB_OK;
}

extern "C" void wb_comm_destroy(wb_comm *cm)
{
  if (!cm)
    return;
  if (cm->ctx)
    cudaSetDevice(cm->ctx->device);
  if (cm->kind==0 && cm->nccl && ncclApi()->lib)
    ncclApi()->CommDestroy(cm->nccl);
  cm->small.release(); cm->snap.release(); cm->tmp.release();
  delete cm;
}

// ============================================================================ the sharded pipeline

struct WbShardMeta                     // what every rank tells every other rank (fixed size: one all-gather)
{
  double bbox[6];                      // min x,y,z, max x,y,z over the rank's header corners
  double prm[4];
  unsigned long long n;
  int nSegs,pad_;
  struct { unsigned long long count; double scale[3],offset[3],unit; } seg[WB_SHARD_MAXSEG];
};

struct WbShard
{
  DevBuf<int> ox,oy,oz;                // the rank's own decoded columns, kept across the two stages
  DevBuf<uint8_t> oc,oret;
  DevBuf<uint32_t> blockCounts,blockOff;
  DevBuf<int4> sendRows,recvRows;
  DevBuf<uint8_t> grid;
  DevBuf<int> ext;
  std::vector<WbShardMeta> meta;
  int maxSegs=1;                       // most input files any rank holds
  wb_shard_stats st{};
};

namespace
{

struct StageClock
{
  wb_ctx *ctx;
  std::chrono::steady_clock::time_point t;
  explicit StageClock(wb_ctx *c):ctx(c),t(std::chrono::steady_clock::now()) {}
  double lap()                         // milliseconds since the last lap, with the stream drained
  {
    cudaStreamSynchronize(ctx->st);
    auto now=std::chrono::steady_clock::now();
    double ms=std::chrono::duration<double,std::milli>(now-t).count();
    t=now;
    return ms;
  }
};

void resetPoints(wb_ctx *ctx)
// forget the points but keep geometry, parameters, tile table and allocations
{
  ctx->n=ctx->nValid=0;
  ctx->nDup=0;
  ctx->segs.clear();
  ctx->recSegs.clear();
  ctx->phase=PH_EMPTY;
  ctx->nPairs=0;
  ctx->nLeaves=0;
}

int haloExchange(wb_ctx *ctx,wb_comm *cm,WbShard &S,const std::vector<WbSegment> &ownSegs,uint64_t nOwn,
                 int mode,double radius,double zTop,double porMax,const std::vector<double> &lo,const std::vector<double> &hi,
                 double ownLo,double ownHi,double reachMax,
                 std::vector<uint64_t> &cntFrom /* [world*MAXSEG]: rows of (sender, sender's segment) */,uint64_t *nRecv,
                 double *msSelect,double *msExchange)
{
  const int W=cm->world,R=cm->rank;
  cudaStream_t st=ctx->st;
  StageClock clk(ctx);
  // destinations that any own point can reach at all
  std::vector<int> dests;
  for (int k=0;k<W;k++)
    if (k!=R && ownHi>=lo[k]-reachMax && ownLo<=hi[k]+reachMax)
      dests.push_back(k);
  uint32_t totalBlocks=0;
  std::vector<uint32_t> blockBase(ownSegs.size()+1,0);
  for (size_t s=0;s<ownSegs.size();s++)
  {
    blockBase[s]=totalBlocks;
    totalBlocks+=(uint32_t)wb_div_up(ownSegs[s].count,WB_HALO_THREADS);
  }
  blockBase[ownSegs.size()]=totalBlocks;
  std::vector<uint64_t> sendCnt((size_t)W*WB_SHARD_MAXSEG,0);      // [dest][own segment]
  std::vector<uint64_t> sOff(W,0),sCnt(W,0);
  uint64_t nSend=0;
  // groups of WB_HALO_MAXDEST destinations (one group in practice: the two neighbours)
  struct Group { size_t first; int n; uint64_t rowBase; };
  std::vector<Group> groups;
  for (size_t g0=0;g0<dests.size();g0+=WB_HALO_MAXDEST)
    groups.push_back(Group{g0,(int)std::min<size_t>(WB_HALO_MAXDEST,dests.size()-g0),0});
  const uint64_t tableLen=(uint64_t)WB_HALO_MAXDEST*totalBlocks+1;
  CK(S.blockCounts.ensure(tableLen*std::max<size_t>(1,groups.size())));
  CK(S.blockOff.ensure(tableLen*std::max<size_t>(1,groups.size())));
  CK(ctx->blockSums.ensure(wb_div_up(tableLen,WB_SCAN_TILE)+1024));
  std::vector<WbHaloArgs> gargs(groups.size());
  std::vector<std::vector<uint32_t>> hOff(groups.size());
  for (size_t g=0;g<groups.size();g++)
  {
    WbHaloArgs &a=gargs[g];
    memset(&a,0,sizeof(a));
    a.nDest=groups[g].n;
    a.mode=mode;
    a.radius=radius;
    a.zTop=zTop;
    a.porMax=porMax;
    a.slope=ctx->prm.maxSlope;
    for (int d=0;d<a.nDest;d++)
    {
      a.lo[d]=lo[dests[groups[g].first+d]];
      a.hi[d]=hi[dests[groups[g].first+d]];
    }
    uint32_t *cnt=S.blockCounts.p+g*tableLen,*off=S.blockOff.p+g*tableLen;
    CK(cudaMemsetAsync(cnt,0,tableLen*sizeof(uint32_t),st));
    for (size_t s=0;s<ownSegs.size();s++)
      if (ownSegs[s].count)
      {
        wb_halo_count_kernel<<<gridFor(ownSegs[s].count,WB_HALO_THREADS),WB_HALO_THREADS,0,st>>>(
            S.ox.p,S.oz.p,S.oret.p,ownSegs[s].first,ownSegs[s].count,ownSegs[s],a,cnt,blockBase[s],totalBlocks);
        ctx->stats.kernel_launches++;
      }
    KCHECK();
    CK(wb_exclusive_scan(cnt,off,tableLen,ctx->blockSums.p,ctx->blockSums.cap,st,&ctx->stats.kernel_launches));
    hOff[g].resize(tableLen);
    CK(cudaMemcpyAsync(hOff[g].data(),off,tableLen*sizeof(uint32_t),cudaMemcpyDeviceToHost,st));
  }
  CK(cudaStreamSynchronize(st));
  for (size_t g=0;g<groups.size();g++)
  {
    groups[g].rowBase=nSend;
    for (int d=0;d<groups[g].n;d++)
    {
      const int k=dests[groups[g].first+d];
      const uint32_t *o=hOff[g].data()+(uint64_t)d*totalBlocks;
      sOff[k]=(nSend+o[0])*sizeof(int4);
      for (size_t s=0;s<ownSegs.size();s++)
      {
        // the table is scanned as one array: the entry after destination d's last block is the next destination's first
        const uint64_t a=o[blockBase[s]],b=o[blockBase[s+1]];
        sendCnt[(size_t)k*WB_SHARD_MAXSEG+s]=b-a;
        sCnt[k]+=(b-a)*sizeof(int4);
      }
    }
    nSend+=hOff[g][(uint64_t)groups[g].n*totalBlocks];
  }
  CK(S.sendRows.ensure(nSend+1));
  for (size_t g=0;g<groups.size();g++)
    for (size_t s=0;s<ownSegs.size();s++)
      if (ownSegs[s].count)
      {
        wb_halo_scatter_kernel<<<gridFor(ownSegs[s].count,WB_HALO_THREADS),WB_HALO_THREADS,0,st>>>(
            S.ox.p,S.oy.p,S.oz.p,S.oc.p,S.oret.p,ownSegs[s].first,ownSegs[s].count,ownSegs[s],(int)s,gargs[g],
            S.blockOff.p+g*tableLen,blockBase[s],totalBlocks,S.sendRows.p+groups[g].rowBase);
        ctx->stats.kernel_launches++;
      }
  KCHECK();
  *msSelect+=clk.lap();
  // who sends what: every rank's [dest][segment] counts to everybody
  // (only the columns any rank uses: the message stays a few hundred bytes for the usual one file per rank — a host
  //  copy of tens of KB goes through the copy engine, behind whatever bulk upload another context has queued there)
  const int MS=S.maxSegs;
  std::vector<uint64_t> mineCnt((size_t)W*MS),allCnt((size_t)W*W*MS);
  for (int k=0;k<W;k++)
    for (int sg=0;sg<MS;sg++)
      mineCnt[(size_t)k*MS+sg]=sendCnt[(size_t)k*WB_SHARD_MAXSEG+sg];
  int rc=commAllGatherHost(cm,mineCnt.data(),allCnt.data(),mineCnt.size()*sizeof(uint64_t));
  if (rc)
    return rc;
  std::vector<uint64_t> rOff(W,0),rCnt(W,0);
  uint64_t total=0;
  cntFrom.assign((size_t)W*WB_SHARD_MAXSEG,0);
  for (int k=0;k<W;k++)
  {
    rOff[k]=total*sizeof(int4);
    for (int s=0;s<MS;s++)
    {
      const uint64_t c=allCnt[((size_t)k*W+R)*MS+s];
      cntFrom[(size_t)k*WB_SHARD_MAXSEG+s]=c;
      rCnt[k]+=c*sizeof(int4);
      total+=c;
    }
  }
  CK(S.recvRows.ensure(total+1));
  if ((rc=commAllToAllV(cm,S.sendRows.p,sOff.data(),sCnt.data(),S.recvRows.p,rOff.data(),rCnt.data())))
    return rc;
  *nRecv=total;
  *msExchange+=clk.lap();
  return WB_OK;
}

int assemble(wb_ctx *ctx,wb_comm *cm,WbShard &S,const std::vector<WbSegment> &ownSegs,uint64_t nOwn,uint64_t ownDropped,
             const std::vector<uint64_t> &cntFrom,uint64_t nRecv)
// local input order = [halo from lower ranks | own | halo from higher ranks]
{
  const int W=cm->world,R=cm->rank;
  cudaStream_t st=ctx->st;
  resetPoints(ctx);
  int rc;
  if ((rc=ensurePointArrays(ctx,nOwn+nRecv)))
    return rc;
  CK(cudaMemcpyAsync(ctx->counters.p+2,&ownDropped,sizeof(unsigned long long),cudaMemcpyHostToDevice,st));
  uint64_t row=0;
  for (int k=0;k<W;k++)
  {
    if (k==R)
    {
      ctx->ownFirst=(uint32_t)ctx->n;
      const uint64_t base=ctx->n;
      if (nOwn)
      {
        wb_copy_columns_kernel<<<gridFor(nOwn,256),256,0,st>>>(S.ox.p,S.oy.p,S.oz.p,S.oc.p,S.oret.p,nOwn,
            ctx->xi.p+base,ctx->yi.p+base,ctx->zi.p+base,ctx->cls.p+base,ctx->ret.p+base);
        ctx->stats.kernel_launches++;
      }
      for (const WbSegment &sg:ownSegs)
        if ((rc=addSegment(ctx,sg.count,sg.scale,sg.offset,sg.unit)))
          return rc;
      ctx->ownEnd=(uint32_t)ctx->n;
      continue;
    }
    const WbShardMeta &m=S.meta[k];
    for (int s=0;s<m.nSegs;s++)
    {
      const uint64_t c=cntFrom[(size_t)k*WB_SHARD_MAXSEG+s];
      if (!c)
        continue;
      const uint64_t base=ctx->n;
      wb_halo_unpack_kernel<<<gridFor(c,256),256,0,st>>>(S.recvRows.p+row,c,ctx->xi.p+base,ctx->yi.p+base,ctx->zi.p+base,
                                                         ctx->cls.p+base,ctx->ret.p+base);
      ctx->stats.kernel_launches++;
      if ((rc=addSegment(ctx,c,m.seg[s].scale,m.seg[s].offset,m.seg[s].unit)))
        return rc;
      row+=c;
    }
  }
  KCHECK();
  if (row!=nRecv)
    return fail(ctx,WB_ERR_STATE,"internal: halo rows %llu of %llu placed",(unsigned long long)row,(unsigned long long)nRecv);
  return WB_OK;
}

} // namespace

extern "C" int wb_shard_run(wb_ctx *ctx,wb_comm *cm)
{
  if (!ctx || !cm || cm->ctx!=ctx)
    return WB_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (ctx->phase!=PH_LOADED || !ctx->n)
    return fail(ctx,WB_ERR_STATE,"wb_shard_run follows wb_add_extent + wb_add_las* of the rank's own files");
  if (ctx->keepRecords)
    return fail(ctx,WB_ERR_STATE,"wb_keep_records is a single-GPU feature");
  if (ctx->segs.size()>WB_SHARD_MAXSEG)
    return fail(ctx,WB_ERR_ARG,"at most %d input files per rank",WB_SHARD_MAXSEG);
  if (ctx->corners.empty())
    return fail(ctx,WB_ERR_STATE,"no extents: call wb_add_extent for the rank's files");
  if (!ctx->shard)
    ctx->shard=new WbShard;
  WbShard &S=*ctx->shard;
  const int W=cm->world,R=cm->rank;
  cudaStream_t st=ctx->st;
  int rc;
  memset(&S.st,0,sizeof(S.st));
  StageClock clk(ctx);
  const uint64_t nOwn=ctx->n;
  const std::vector<WbSegment> ownSegs=ctx->segs;
  // ---- own columns aside
  CK(S.ox.ensure(nOwn)); CK(S.oy.ensure(nOwn)); CK(S.oz.ensure(nOwn)); CK(S.oc.ensure(nOwn)); CK(S.oret.ensure(nOwn));
  wb_copy_columns_kernel<<<gridFor(nOwn,256),256,0,st>>>(ctx->xi.p,ctx->yi.p,ctx->zi.p,ctx->cls.p,ctx->ret.p,nOwn,
                                                        S.ox.p,S.oy.p,S.oz.p,S.oc.p,S.oret.p);
  ctx->stats.kernel_launches++;
  KCHECK();
  unsigned long long ownDropped=0;
  CK(cudaMemcpyAsync(&ownDropped,ctx->counters.p+2,sizeof(ownDropped),cudaMemcpyDeviceToHost,st));
  // ---- everybody's extents, parameters and segments
  WbShardMeta mine;
  memset(&mine,0,sizeof(mine));
  for (int k=0;k<3;k++)
  {
    mine.bbox[k]=INFINITY;
    mine.bbox[3+k]=-INFINITY;
  }
  for (size_t i=0;i+2<ctx->corners.size();i+=3)
    for (int k=0;k<3;k++)
    {
      mine.bbox[k]=std::min(mine.bbox[k],ctx->corners[i+k]);
      mine.bbox[3+k]=std::max(mine.bbox[3+k],ctx->corners[i+k]);
    }
  mine.prm[0]=ctx->prm.tileSize; mine.prm[1]=ctx->prm.maxSlope; mine.prm[2]=ctx->prm.thickness; mine.prm[3]=ctx->prm.minHyp;
  mine.n=nOwn;
  mine.nSegs=(int)ownSegs.size();
  for (size_t s=0;s<ownSegs.size();s++)
  {
    mine.seg[s].count=ownSegs[s].count;
    for (int k=0;k<3;k++)
    {
      mine.seg[s].scale[k]=ownSegs[s].scale[k];
      mine.seg[s].offset[k]=ownSegs[s].offset[k];
    }
    mine.seg[s].unit=ownSegs[s].unit;
  }
  S.meta.assign(W,WbShardMeta());
  {
    // the fixed part first, then the per-file headers padded to the largest file count of any rank
    const size_t headBytes=offsetof(WbShardMeta,seg);
    std::vector<uint8_t> heads(headBytes*W);
    if ((rc=commAllGatherHost(cm,&mine,heads.data(),headBytes)))
      return rc;
    int maxSegs=1;
    for (int k=0;k<W;k++)
    {
      memcpy(&S.meta[k],heads.data()+headBytes*k,headBytes);
      if (S.meta[k].nSegs<0 || S.meta[k].nSegs>WB_SHARD_MAXSEG)
        return fail(ctx,WB_ERR_ARG,"rank %d reports %d input files",k,S.meta[k].nSegs);
      maxSegs=std::max(maxSegs,S.meta[k].nSegs);
    }
    S.maxSegs=maxSegs;
    const size_t segBytes=sizeof(mine.seg[0])*(size_t)maxSegs;
    std::vector<uint8_t> segs(segBytes*W);
    if ((rc=commAllGatherHost(cm,mine.seg,segs.data(),segBytes)))
      return rc;
    for (int k=0;k<W;k++)
      memcpy(S.meta[k].seg,segs.data()+segBytes*k,segBytes);
  }
  std::vector<double> lo(W),hi(W),cuts(W+1);
  double zmax=-INFINITY;
  {
    std::vector<double> corners;
    for (int k=0;k<W;k++)
    {
      const WbShardMeta &m=S.meta[k];
      if (memcmp(m.prm,mine.prm,sizeof(mine.prm)))
        return fail(ctx,WB_ERR_ARG,"rank %d runs with other parameters (tileSize, maxSlope, thickness, minHyperboloidSize)",k);
      if (!m.n)
        return fail(ctx,WB_ERR_ARG,"rank %d holds no points",k);
      for (int j=0;j<6;j++)
        corners.push_back(m.bbox[j]);
      lo[k]=m.bbox[0];
      hi[k]=m.bbox[3];
      zmax=std::max(zmax,m.bbox[5]);
      if (k && (lo[k]<lo[k-1] || hi[k]<hi[k-1]))
        return fail(ctx,WB_ERR_ARG,"ranks must hold x-strips in ascending order (rank %d starts at %g, rank %d at %g)",
                    k-1,lo[k-1],k,lo[k]);
    }
    wb_geometry &g=ctx->geom;
    wbhost::sizeFit(corners.data(),2*W,g.root_center,&g.root_side);
    wbhost::bboxCube(corners.data(),2*W,g.cube);
    ctx->geomOverride=true;
    if ((rc=computeGeometry(ctx)))
      return rc;
  }
  cuts[0]=-INFINITY;
  cuts[W]=INFINITY;
  for (int k=1;k<W;k++)
    cuts[k]=0.5*(hi[k-1]+lo[k]);                 // ownership of tile centres: the middle of the gap between strips
  const double ownLo=cuts[R],ownHi=cuts[R+1];
  const double spacing=ctx->geom.spacing;
  S.st.ms_setup=clk.lap();
  // ---- stage A: narrow halo, build, scan
  std::vector<uint64_t> cntFrom;
  uint64_t nRecv=0;
  const double r1=WB_SHARD_SCAN_HALO*spacing;
  if ((rc=haloExchange(ctx,cm,S,ownSegs,nOwn,1,r1,0,0,lo,hi,lo[R],hi[R],r1,cntFrom,&nRecv,&S.st.ms_select,&S.st.ms_exchange)))
    return rc;
  S.st.n_halo_scan=nRecv;
  if ((rc=assemble(ctx,cm,S,ownSegs,nOwn,ownDropped,cntFrom,nRecv)))
    return rc;
  if ((rc=wb_build(ctx)))
    return rc;
  S.st.ms_build_scan=clk.lap();
  if ((rc=wb_scan(ctx)))
    return rc;
  S.st.ms_scan=clk.lap();
  const uint32_t nList=(uint32_t)ctx->stats.n_tiles_nonempty;
  // ---- the populated / tree grid over the lattice box of ALL ranks' owned tiles
  {
    const int init[4]={INT_MAX,INT_MAX,INT_MIN,INT_MIN};
    int ext[4];
    CK(S.ext.ensure(4));
    CK(cudaMemcpyAsync(S.ext.p,init,sizeof(init),cudaMemcpyHostToDevice,st));
    if (nList)
    {
      wb_tile_extent_list_kernel<<<gridFor(nList,256),256,0,st>>>(ctx->tileList.p,nList,ctx->snake,ownLo,ownHi,S.ext.p);
      ctx->stats.kernel_launches++;
    }
    KCHECK();
    CK(cudaMemcpyAsync(ext,S.ext.p,sizeof(ext),cudaMemcpyDeviceToHost,st));
    CK(cudaStreamSynchronize(st));
    std::vector<int> all(4*W);
    if ((rc=commAllGatherHost(cm,ext,all.data(),sizeof(ext))))
      return rc;
    for (int k=0;k<W;k++)
    {
      ext[0]=std::min(ext[0],all[4*k]);
      ext[1]=std::min(ext[1],all[4*k+1]);
      ext[2]=std::max(ext[2],all[4*k+2]);
      ext[3]=std::max(ext[3],all[4*k+3]);
    }
    uint64_t cells=1;
    if (ext[0]<=ext[2])
      cells=(uint64_t)((long long)ext[2]-ext[0]+1)*(uint64_t)((long long)ext[3]-ext[1]+1);
    else
      ext[0]=ext[1]=ext[2]=ext[3]=0;
    CK(cudaMemcpyAsync(S.ext.p,ext,sizeof(ext),cudaMemcpyHostToDevice,st));
    CK(S.grid.ensure(cells+16));
    CK(cudaMemsetAsync(S.grid.p,0,cells,st));
    if (nList)
    {
      wb_tile_grid_list_kernel<<<gridFor(nList,256),256,0,st>>>(ctx->tileList.p,nList,ctx->tTree.p,ctx->snake,ownLo,ownHi,
                                                               S.ext.p,S.grid.p);
      ctx->stats.kernel_launches++;
    }
    KCHECK();
    if ((rc=commAllReduceMaxU8(cm,S.grid.p,cells)))
      return rc;
    S.st.grid_cells=cells;
    S.st.ms_grid=clk.lap();
    // ---- postscan of every tile this rank scanned (the ones its own points use are among them)
    if (nList)
    {
      wb_postscan_list_kernel<<<gridFor(nList,128),128,0,st>>>(ctx->tileList.p,nList,ctx->tTree.p,ctx->snake,S.ext.p,S.grid.p,
                                                              ctx->tHyp.p);
      ctx->stats.kernel_launches++;
    }
    KCHECK();
  }
  // ---- how far classify can reach: the largest hyperboloid anywhere
  double porMax=0;
  {
    CK(cudaMemsetAsync(ctx->counters.p+7,0,sizeof(unsigned long long),st));
    if (nList)
    {
      wb_max_hyp_list_kernel<<<148*4,256,0,st>>>(ctx->tileList.p,nList,ctx->tHyp.p,ctx->snake,ownLo,ownHi,ctx->counters.p+7);
      ctx->stats.kernel_launches++;
    }
    KCHECK();
    unsigned long long bits=0;
    CK(cudaMemcpyAsync(&bits,ctx->counters.p+7,sizeof(bits),cudaMemcpyDeviceToHost,st));
    CK(cudaStreamSynchronize(st));
    std::vector<unsigned long long> all(W);
    if ((rc=commAllGatherHost(cm,&bits,all.data(),sizeof(bits))))
      return rc;
    for (int k=0;k<W;k++)
      bits=std::max(bits,all[k]);
    double h;
    memcpy(&h,&bits,sizeof(h));
    porMax=h*ctx->prm.maxSlope*ctx->prm.maxSlope;
  }
  S.st.por_max=porMax;
  S.st.ms_postscan=clk.lap();
  // ---- stage B: wide halo, rebuild, classify own points
  {
    const double zTop=zmax-ctx->prm.thickness;
    double zminAll=INFINITY;
    for (int k=0;k<W;k++)
      zminAll=std::min(zminAll,S.meta[k].bbox[2]);
    const double dzMax=std::max(zTop-zminAll,0.0);
    const double reachMax=sqrt(dzMax*dzMax+2.0*porMax*dzMax)/ctx->prm.maxSlope*(1+1e-9)+1e-6;
    if ((rc=haloExchange(ctx,cm,S,ownSegs,nOwn,2,0,zTop,porMax,lo,hi,lo[R],hi[R],reachMax,cntFrom,&nRecv,
                         &S.st.ms_select,&S.st.ms_exchange)))
      return rc;
  }
  S.st.n_halo_classify=nRecv;
  if ((rc=assemble(ctx,cm,S,ownSegs,nOwn,ownDropped,cntFrom,nRecv)))
    return rc;
  // classify walks the store along a Hilbert curve over xy whatever order it is kept in, and nothing else is asked
  // of this second store: it is built in that order straight away (no Morton sort, no octree leaves)
  if ((rc=buildStore(ctx,true)))
    return rc;
  S.st.ms_build_classify=clk.lap();
  if ((rc=wb_assign(ctx)))
    return rc;
  ctx->phase=PH_POSTSCANNED;
  S.st.ms_assign=clk.lap();
  if ((rc=wb_classify(ctx)))
    return rc;
  S.st.ms_classify=clk.lap();
  S.st.n_own=nOwn;
  S.st.own_first=ctx->ownFirst;
  S.st.bytes_sent=cm->bytesSent;
  S.st.bytes_received=cm->bytesReceived;
  cm->bytesSent=cm->bytesReceived=0;
  return WB_OK;
}

extern "C" int wb_shard_get_labels(wb_ctx *ctx,uint8_t *labels)
// class byte of the rank's OWN records, in the order they were added
{
  if (!ctx || !labels)
    return WB_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (ctx->phase<PH_CLASSIFIED || !ctx->shard)
    return fail(ctx,WB_ERR_STATE,"not classified by wb_shard_run");
  CK(cudaEventRecord(ctx->evA,ctx->st));
  CK(cudaMemcpyAsync(labels,ctx->labelIn.p+ctx->ownFirst,ctx->ownEnd-ctx->ownFirst,cudaMemcpyDeviceToHost,ctx->st));
  CK(cudaEventRecord(ctx->evB,ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  ctx->stats.ms_d2h=elapsed(ctx->evA,ctx->evB);
  return WB_OK;
}

extern "C" int wb_shard_get_stats(wb_ctx *ctx,wb_shard_stats *out)
{
  if (!ctx || !out)
    return WB_ERR_ARG;
  if (!ctx->shard)
    return fail(ctx,WB_ERR_STATE,"wb_shard_run has not run");
  *out=ctx->shard->st;
  return WB_OK;
}

static void wbShardFree(wb_ctx *ctx)
{
  if (!ctx->shard)
    return;
  WbShard &S=*ctx->shard;
  S.ox.release(); S.oy.release(); S.oz.release(); S.oc.release(); S.oret.release();
  S.blockCounts.release(); S.blockOff.release(); S.sendRows.release(); S.recvRows.release();
  S.grid.release(); S.ext.release();
  delete ctx->shard;
  ctx->shard=nullptr;
}

extern "C" int wb_set_labels(wb_ctx *ctx,const uint8_t *labels)
// Class bytes computed elsewhere (the ranks of a sharded run) into a built store, input order: count, encode and
// write then work as after wb_classify.
{
  if (!ctx || !labels)
    return WB_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (ctx->phase<PH_BUILT)
    return fail(ctx,WB_ERR_STATE,"not built");
  cudaStream_t st=ctx->st;
  CK(cudaMemcpyAsync(ctx->labelIn.p,labels,ctx->n,cudaMemcpyHostToDevice,st));
  wb_sorted_labels_kernel<<<gridFor(ctx->nValid,256),256,0,st>>>(ctx->labelIn.p,ctx->perm,ctx->nValid,ctx->labelSorted.p);
  ctx->stats.kernel_launches++;
  KCHECK();
  CK(cudaStreamSynchronize(st));
  ctx->phase=PH_CLASSIFIED;
  return WB_OK;
}
