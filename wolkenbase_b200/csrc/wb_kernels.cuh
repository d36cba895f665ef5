// wb_kernels.cuh — the sm_100a kernels of the ground-extraction path.
//
// Data layout in HBM (N = points, all SoA, index = input order unless "sorted"):
//   records  u8[N*L]        packed LAS records as read from the file            (las.cpp:735-820)
//   xi,yi,zi i32[N]         decoded integer coordinates; cls u8[N]; ret u8[N]
//   key      u64[N]         21-level Morton key of the reference's ">=center" descent
//   perm     u32[N]         sorted position -> input index (canonical order = key, then input index)
//   sx,sy,sz f64[N]         coordinates in canonical order, exactly (offset+scale*int)*unit
//   bound    {xmin,xmax,ymin,ymax,zmin} f64 per 32-point bucket chunk, and per node of the
//            32-ary hierarchy above the chunks (level l node j covers chunks [j*32^l,(j+1)*32^l))
//   tiles    dense arrays indexed by flowsnake sequence number - lo
#pragma once
#include "wb_device.cuh"
#include <cstdlib>
#if defined(__CUDACC__) && !defined(WB_DEC_NO_BULK)
#include <cuda/barrier>                 // cp.async.bulk (TMA, 1-D) for the decode kernel's staging
#define WB_DEC_BULK 1
#else
#define WB_DEC_BULK 0                   // the host build of tests/simt takes the vector-load path
#endif

#define WB_FULL 0xffffffffu
#define WB_HILBERT_BITS 20

__device__ __forceinline__ unsigned long long wb_hilbert_index(double px,double py,double x0,double y0,double cellsPerUnit)
// index of (px,py) along a Hilbert curve of 2^20 x 2^20 cells whose corner is (x0,y0)
{
  const double lim=(double)((1u<<WB_HILBERT_BITS)-1);
  uint32_t x=(uint32_t)fmin(fmax((px-x0)*cellsPerUnit,0.0),lim);
  uint32_t y=(uint32_t)fmin(fmax((py-y0)*cellsPerUnit,0.0),lim);
  unsigned long long d=0;
  #pragma unroll 4
  for (uint32_t s=1u<<(WB_HILBERT_BITS-1);s;s>>=1)
  {
    const uint32_t rx=(x&s)?1u:0u,ry=(y&s)?1u:0u;
    d=(d<<2)|((3u*rx)^ry);
    if (!ry)
    {
      if (rx)
      {
        x=~x;
        y=~y;
      }
      const uint32_t t=x;
      x=y;
      y=t;
    }
  }
  return d;
}

struct WbBound { double xmin,xmax,ymin,ymax,zmin; };

struct WbSegment            // one wb_add_las call
{
  unsigned long long first,count;
  double scale[3],offset[3],unit;
};
#ifndef WB_MAX_SEGMENTS
#define WB_MAX_SEGMENTS 2048
#endif
struct WbSegments
{
  int n;
  WbSegment s[WB_MAX_SEGMENTS];
};

// ============================================================================ K1: LAS decode
// One CTA decodes 256 consecutive records.  The CTA's byte span is staged in shared memory with
// 16-byte loads (coalesced whatever the record length: 20..38 bytes), fields are pulled out of
// shared memory with unaligned 32-bit extraction, and the SoA columns are stored coalesced.

#define WB_DEC_THREADS 256
#define WB_DEC_MAXLEN 40

__device__ __forceinline__ uint32_t wb_ld32(const uint32_t *w,uint32_t byteOff)
{
  uint32_t i=byteOff>>2,sh=(byteOff&3)*8;
  uint32_t lo=w[i],hi=w[i+1];
  return __funnelshift_r(lo,hi,sh);
}

__global__ void __launch_bounds__(WB_DEC_THREADS)
wb_decode_kernel(const uint8_t *__restrict__ recs,unsigned long long n,int fmt,int recLen,int dropZeros,
                 int *__restrict__ xi,int *__restrict__ yi,int *__restrict__ zi,
                 uint8_t *__restrict__ cls,uint8_t *__restrict__ ret,unsigned long long *nDropped)
{
  __align__(16) __shared__ uint32_t sw[(WB_DEC_THREADS*WB_DEC_MAXLEN+32)/4+4];
  const unsigned long long first=(unsigned long long)blockIdx.x*WB_DEC_THREADS;
  const unsigned long long cnt=n-first<WB_DEC_THREADS?n-first:WB_DEC_THREADS;
  const unsigned long long b0=first*recLen,b1=(first+cnt)*recLen;
  const uintptr_t p0=(uintptr_t)recs+b0;
  const uintptr_t a0=p0&~(uintptr_t)15;               // aligned start (may precede our span)
  const uint32_t head=(uint32_t)(p0-a0);
  const uint32_t total=head+(uint32_t)(b1-b0);
  const uint32_t nvec=(total+15)/16;
  const uintptr_t endAll=(uintptr_t)recs+n*(unsigned long long)recLen;
#if WB_DEC_BULK
  // The CTA's span, rounded out to 16-byte boundaries, is ONE bulk copy into shared memory (cp.async.bulk, the 1-D
  // form of TMA: SASS UBLKCP): a single thread issues it, the bytes arrive by the async proxy and complete an
  // mbarrier transaction — no thread spends registers or load slots on the 5-10 KB.  Only the first and last CTA of
  // a buffer (whose rounded span would reach outside it) take the vector-load path below.
  #pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ cuda::barrier<cuda::thread_scope_block> bar;
  const bool bulk=a0>=(uintptr_t)recs && a0+(uintptr_t)nvec*16<=endAll;       // CTA-uniform
  if (bulk)
  {
    namespace cde=cuda::device::experimental;
    if (threadIdx.x==0)
    {
      init(&bar,WB_DEC_THREADS);
      cde::fence_proxy_async_shared_cta();
    }
    __syncthreads();
    cuda::barrier<cuda::thread_scope_block>::arrival_token token;
    if (threadIdx.x==0)
    {
      cde::cp_async_bulk_global_to_shared(sw,reinterpret_cast<const void *>(a0),nvec*16,bar);
      token=cuda::device::barrier_arrive_tx(bar,1,nvec*16);
    }
    else
      token=bar.arrive();
    bar.wait(std::move(token));
  }
  else
#endif
  for (uint32_t v=threadIdx.x;v<nvec;v+=WB_DEC_THREADS)
  {
    uintptr_t a=a0+(uintptr_t)v*16;
    uint4 q;
    if (a>=(uintptr_t)recs && a+16<=endAll)
      q=*reinterpret_cast<const uint4 *>(a);
    else
    {
      // first/last vector of the whole buffer: never touch bytes outside [recs,endAll)
      uint32_t wv[4]={0,0,0,0};
      for (int b=0;b<16;b++)
      {
        uintptr_t ab=a+b;
        if (ab>=(uintptr_t)recs && ab<endAll)
          wv[b>>2]|=(uint32_t)(*reinterpret_cast<const uint8_t *>(ab))<<((b&3)*8);
      }
      q=make_uint4(wv[0],wv[1],wv[2],wv[3]);
    }
    reinterpret_cast<uint4 *>(sw)[v]=q;
  }
  __syncthreads();
  if (threadIdx.x<cnt)
  {
    uint32_t o=head+threadIdx.x*recLen;
    int x=(int)wb_ld32(sw,o),y=(int)wb_ld32(sw,o+4),z=(int)wb_ld32(sw,o+8);
    uint32_t f=wb_ld32(sw,o+14);                      // bytes 14,15,16,17
    uint8_t r,c;
    if (fmt<6)
    {
      r=f&7;                                          // las.cpp:752
      c=(f>>8)&31;                                    // las.cpp:756-758
    }
    else
    {
      r=f&15;                                         // las.cpp:766
      c=(f>>16)&255;                                  // las.cpp:773
    }
    unsigned long long i=first+threadIdx.x;
    xi[i]=x; yi[i]=y; zi[i]=z;
    cls[i]=c;
    if (r==0 && !dropZeros)
      r=1;                                            // threads.cpp:527-528
    ret[i]=r;
    if (r==0)
      atomicAdd(nDropped,1ull);
  }
}

// ============================================================================ K1b: x-window of a rank
// A rank of a sharded run may be handed WHOLE files and an x-interval (wb_set_window): of every file it keeps the
// records whose x lies in [lo,hi), in file order — so that one big file can be classified by several GPUs without
// being cut up first (each rank decodes all of it; decode is a millisecond per 100 M records).

__global__ void __launch_bounds__(256)
wb_window_flag_kernel(const int *__restrict__ xi,const uint8_t *__restrict__ ret,unsigned long long first,unsigned long long n,
                      WbSegment seg,double lo,double hi,uint32_t *__restrict__ flag,unsigned long long *__restrict__ nDropped)
// flag[t] = record first+t is inside; the dropped-record counter loses the dropped records that go away
{
  unsigned long long t=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  bool inside=false,gone=false;
  if (t<n)
  {
    const double x=wb_coord(seg.offset[0],seg.scale[0],xi[first+t],seg.unit);
    inside=x>=lo && x<hi;
    gone=!inside && ret[first+t]==0;
    flag[t]=inside?1u:0u;
  }
  const unsigned g=__ballot_sync(WB_FULL,gone);
  if ((threadIdx.x&31)==0 && g)
    atomicAdd(nDropped,(unsigned long long)(0ull-(unsigned long long)__popc(g)));
}

__global__ void __launch_bounds__(256)
wb_window_scatter_kernel(const uint32_t *__restrict__ flag,const uint32_t *__restrict__ off,unsigned long long first,
                         unsigned long long n,const int *__restrict__ xi,const int *__restrict__ yi,const int *__restrict__ zi,
                         const uint8_t *__restrict__ cls,const uint8_t *__restrict__ ret,
                         int *__restrict__ ox,int *__restrict__ oy,int *__restrict__ oz,uint8_t *__restrict__ oc,
                         uint8_t *__restrict__ oret)
{
  unsigned long long t=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  if (t>=n || !flag[t])
    return;
  const uint32_t o=off[t];
  ox[o]=xi[first+t]; oy[o]=yi[first+t]; oz[o]=zi[first+t];
  oc[o]=cls[first+t];
  oret[o]=ret[first+t];
}

__global__ void __launch_bounds__(256)
wb_window_copy_back_kernel(const int *__restrict__ ix,const int *__restrict__ iy,const int *__restrict__ iz,
                           const uint8_t *__restrict__ ic,const uint8_t *__restrict__ iret,unsigned long long m,
                           unsigned long long first,int *__restrict__ xi,int *__restrict__ yi,int *__restrict__ zi,
                           uint8_t *__restrict__ cls,uint8_t *__restrict__ ret)
{
  unsigned long long t=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  if (t>=m)
    return;
  xi[first+t]=ix[t]; yi[first+t]=iy[t]; zi[first+t]=iz[t];
  cls[first+t]=ic[t];
  ret[first+t]=iret[t];
}

// ============================================================================ K2: Morton keys

__device__ __forceinline__ uint32_t wb_cell21(double v,double c,double side)
// Index (21 bits) of the cell of width side/2^21 that Octree::findBlock's descent
// (octree.cpp:199-216: bit = coordinate >= centre, 21 times) ends in.  The descent is a binary
// search for the cell [lo,lo+w) containing v, with every boundary a dyadic number that is exact in
// FP64; so the cell is found by one scaled subtraction and then CHECKED against the exact
// boundaries (and moved by one if the rounded quotient landed next door).  Coordinates outside
// the root cube clamp to the first/last cell, as the comparison chain does.
{
  const double w=side*(1.0/2097152.0),lo0=c-0.5*side;     // exact: side is a power of two
  double q=floor((v-lo0)/w);
  q=fmin(fmax(q,0.0),2097151.0);
  double lo=lo0+q*w;                                       // exact (dyadic, few significant bits)
  if (v<lo && q>0.0)
    q-=1.0;
  else if (v>=lo+w && q<2097151.0)
    q+=1.0;
  return (uint32_t)q;
}

__device__ __forceinline__ unsigned long long wb_spread3(uint32_t v)
// 21 bits -> every third bit of 63
{
  unsigned long long x=v&0x1fffffull;
  x=(x|(x<<32))&0x1f00000000ffffull;
  x=(x|(x<<16))&0x1f0000ff0000ffull;
  x=(x|(x<<8))&0x100f00f00f00f00full;
  x=(x|(x<<4))&0x10c30c30c30c30c3ull;
  x=(x|(x<<2))&0x1249249249249249ull;
  return x;
}

__device__ __forceinline__ unsigned long long wb_morton(double x,double y,double z,
                                                        double cx,double cy,double cz,double side)
// 63-bit key: per level (most significant first) the child index z*4+y*2+x of the descent.
{
  return wb_spread3(wb_cell21(x,cx,side))|(wb_spread3(wb_cell21(y,cy,side))<<1)|(wb_spread3(wb_cell21(z,cz,side))<<2);
}

__global__ void __launch_bounds__(256)
wb_keygen_kernel(const int *__restrict__ xi,const int *__restrict__ yi,const int *__restrict__ zi,
                 const uint8_t *__restrict__ ret,unsigned long long first,unsigned long long cnt,
                 WbSegment seg,int segIndex,double cx,double cy,double cz,double side,int hilbert,
                 unsigned long long *__restrict__ key,uint32_t *__restrict__ idx,int4 *__restrict__ packed)
// hilbert: the key of a CLASSIFY-ONLY store (wb_hilbert_index over xy, see "classify order" below) instead of the
// octree's Morton key
{
  unsigned long long t=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  if (t>=cnt)
    return;
  unsigned long long i=first+t;
  double x=wb_coord(seg.offset[0],seg.scale[0],xi[i],seg.unit);
  double y=wb_coord(seg.offset[1],seg.scale[1],yi[i],seg.unit);
  double z=wb_coord(seg.offset[2],seg.scale[2],zi[i],seg.unit);
  unsigned long long k=hilbert?wb_hilbert_index(x,y,cx-side,cy-side,(double)(1u<<WB_HILBERT_BITS)/(2*side))
                              :wb_morton(x,y,z,cx,cy,cz,side);
  if (ret[i]==0)
    k=~0ull;                                          // dropped record: sorts behind every real key
  key[i]=k;
  idx[i]=(uint32_t)i;
  packed[i]=make_int4(xi[i],yi[i],zi[i],segIndex);   // what the gather to sorted order reads: ONE 16-byte element per point
}

// ============================================================================ K3: gather to canonical order

__global__ void __launch_bounds__(256)
wb_gather_kernel(const uint32_t *__restrict__ perm,unsigned long long n,const int4 *__restrict__ packed,
                 const WbSegments *__restrict__ segs,
                 double *__restrict__ sx,double *__restrict__ sy,double *__restrict__ sz)
// The permuted read is the expensive part (every access its own 32-byte sector): the three integers and the index of
// their file's header come as one int4 written by wb_keygen_kernel, not as three 4-byte reads from three columns plus
// a search for the file (51 GB of DRAM traffic for 125 M points before, ncu; a third of that now).
{
  unsigned long long j=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  if (j>=n)
    return;
  const int4 q=packed[perm[j]];
  const WbSegment &s=segs->s[q.w];
  sx[j]=wb_coord(s.offset[0],s.scale[0],q.x,s.unit);
  sy[j]=wb_coord(s.offset[1],s.scale[1],q.y,s.unit);
  sz[j]=wb_coord(s.offset[2],s.scale[2],q.z,s.unit);
}

// ============================================================================ K4: leaf split
// Level-synchronous top-down bucket split.  A node (a run of the sorted keys sharing a prefix of
// 3*depth bits) with more than 537 points is internal (OctStore::put/split, octree.cpp:849-876,
// 1295-1338); its non-empty children are found by binary search for the 7 octant boundaries.
// Children with <= 537 points are leaves: the depth is written at the leaf's first point.

struct WbNode { unsigned long long first; uint32_t count; };

__global__ void __launch_bounds__(128)
wb_split_kernel(const unsigned long long *__restrict__ keys,const WbNode *__restrict__ nodes,uint32_t nNodes,
                int depth /* of these nodes */,WbNode *__restrict__ next,uint32_t *__restrict__ nNext,
                uint8_t *__restrict__ leafDepth,uint32_t *__restrict__ nLeaves)
{
  uint32_t t=blockIdx.x*blockDim.x+threadIdx.x;
  if (t>=nNodes*8)
    return;
  WbNode nd=nodes[t>>3];
  const int ch=t&7,shift=3*(20-depth);
  // first index in [first,first+count) whose digit >= ch, and >= ch+1
  unsigned long long lo=nd.first,hi=nd.first+nd.count,a,b;
  {
    unsigned long long l=lo,h=hi;
    while (l<h)
    {
      unsigned long long m=(l+h)>>1;
      if (((keys[m]>>shift)&7)<(unsigned)ch) l=m+1; else h=m;
    }
    a=l;
    h=hi;
    while (l<h)
    {
      unsigned long long m=(l+h)>>1;
      if (((keys[m]>>shift)&7)<=(unsigned)ch) l=m+1; else h=m;
    }
    b=l;
  }
  if (b==a)
    return;
  uint32_t c=(uint32_t)(b-a);
  if (c<=537 || depth+1>=21)
  {
    leafDepth[a]=(uint8_t)(depth+1);
    atomicAdd(nLeaves,1u);
  }
  else
  {
    uint32_t s=atomicAdd(nNext,1u);
    next[s].first=a;
    next[s].count=c;
  }
}

__global__ void __launch_bounds__(256)
wb_leaf_flag_kernel(const uint8_t *__restrict__ leafDepth,unsigned long long n,uint32_t *__restrict__ flag)
{
  unsigned long long j=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  if (j<n)
    flag[j]=leafDepth[j]!=0;
}

struct WbLeafDev { unsigned long long first; uint32_t count; int depth; double low,high; };

__global__ void __launch_bounds__(256)
wb_leaf_emit_kernel(const uint8_t *__restrict__ leafDepth,const uint32_t *__restrict__ pos,unsigned long long n,
                    WbLeafDev *__restrict__ leaves)
{
  unsigned long long j=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  if (j<n && leafDepth[j])
  {
    leaves[pos[j]].first=j;
    leaves[pos[j]].depth=leafDepth[j];
  }
}

__global__ void __launch_bounds__(256)
wb_leaf_finish_kernel(WbLeafDev *__restrict__ leaves,uint32_t nLeaves,unsigned long long n,
                      const double *__restrict__ sz)
// one warp per leaf: count from the next leaf's start, z range of the bucket (OctBuffer low/high)
{
  uint32_t w=(blockIdx.x*blockDim.x+threadIdx.x)>>5,lane=threadIdx.x&31;
  if (w>=nLeaves)
    return;
  unsigned long long a=leaves[w].first,b=w+1<nLeaves?leaves[w+1].first:n;
  double lo=INFINITY,hi=-INFINITY;
  for (unsigned long long j=a+lane;j<b;j+=32)
  {
    double z=sz[j];
    lo=fmin(lo,z);
    hi=fmax(hi,z);
  }
  #pragma unroll
  for (int o=16;o;o>>=1)
  {
    lo=fmin(lo,__shfl_xor_sync(WB_FULL,lo,o));
    hi=fmax(hi,__shfl_xor_sync(WB_FULL,hi,o));
  }
  if (lane==0)
  {
    leaves[w].count=(uint32_t)(b-a);
    leaves[w].low=lo;
    leaves[w].high=hi;
  }
}

// ============================================================================ K5: bucket hierarchy

__global__ void __launch_bounds__(256)
wb_chunk_bounds_kernel(const double *__restrict__ sx,const double *__restrict__ sy,const double *__restrict__ sz,
                       unsigned long long n,WbBound *__restrict__ out,uint32_t nChunks)
{
  uint32_t w=(blockIdx.x*blockDim.x+threadIdx.x)>>5,lane=threadIdx.x&31;
  if (w>=nChunks)
    return;
  unsigned long long j=(unsigned long long)w*32+lane;
  double x0=INFINITY,x1=-INFINITY,y0=INFINITY,y1=-INFINITY,z0=INFINITY;
  if (j<n)
  {
    x0=x1=sx[j];
    y0=y1=sy[j];
    z0=sz[j];
  }
  #pragma unroll
  for (int o=16;o;o>>=1)
  {
    x0=fmin(x0,__shfl_xor_sync(WB_FULL,x0,o));
    x1=fmax(x1,__shfl_xor_sync(WB_FULL,x1,o));
    y0=fmin(y0,__shfl_xor_sync(WB_FULL,y0,o));
    y1=fmax(y1,__shfl_xor_sync(WB_FULL,y1,o));
    z0=fmin(z0,__shfl_xor_sync(WB_FULL,z0,o));
  }
  if (lane==0)
  {
    WbBound b={x0,x1,y0,y1,z0};
    out[w]=b;
  }
}

__global__ void __launch_bounds__(256)
wb_node_bounds_kernel(const WbBound *__restrict__ child,uint32_t nChild,WbBound *__restrict__ out,uint32_t nOut)
{
  uint32_t w=(blockIdx.x*blockDim.x+threadIdx.x)>>5,lane=threadIdx.x&31;
  if (w>=nOut)
    return;
  uint32_t j=w*32+lane;
  double x0=INFINITY,x1=-INFINITY,y0=INFINITY,y1=-INFINITY,z0=INFINITY;
  if (j<nChild)
  {
    WbBound b=child[j];
    x0=b.xmin; x1=b.xmax; y0=b.ymin; y1=b.ymax; z0=b.zmin;
  }
  #pragma unroll
  for (int o=16;o;o>>=1)
  {
    x0=fmin(x0,__shfl_xor_sync(WB_FULL,x0,o));
    x1=fmax(x1,__shfl_xor_sync(WB_FULL,x1,o));
    y0=fmin(y0,__shfl_xor_sync(WB_FULL,y0,o));
    y1=fmax(y1,__shfl_xor_sync(WB_FULL,y1,o));
    z0=fmin(z0,__shfl_xor_sync(WB_FULL,z0,o));
  }
  if (lane==0)
  {
    WbBound b={x0,x1,y0,y1,z0};
    out[w]=b;
  }
}

// ============================================================================ K6: tile membership

__global__ void __launch_bounds__(128)
wb_member_count_kernel(const double *__restrict__ sx,const double *__restrict__ sy,unsigned long long n,
                       WbSnake snake,uint32_t *__restrict__ cnt,uint4 *__restrict__ tilesOf,
                       uint32_t *__restrict__ winner)
// For every point (canonical order): the tiles whose cylinder contains it (<=3), and the LAST
// of them in flowsnake order — the one whose classifyCylinder call writes the surviving label
// in a 1-thread reference run (classify.cpp:158-165).
{
  unsigned long long k=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  if (k>=n)
    return;
  int t[4]={-1,-1,-1,-1};
  int c=wb_covering_tiles(snake,sx[k],sy[k],t);
  int best=-1;
  for (int j=0;j<c;j++)
    best=max(best,t[j]);
  cnt[k]=c;
  tilesOf[k]=make_uint4((uint32_t)t[0],(uint32_t)t[1],(uint32_t)t[2],(uint32_t)t[3]);
  winner[k]=(uint32_t)best;
}

__global__ void __launch_bounds__(256)
wb_member_fill_kernel(const uint32_t *__restrict__ cnt,const uint32_t *__restrict__ off,
                      const uint4 *__restrict__ tilesOf,unsigned long long n,
                      uint32_t *__restrict__ pairKey,uint32_t *__restrict__ pairVal)
{
  unsigned long long k=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  if (k>=n)
    return;
  uint32_t c=cnt[k],o=off[k];
  uint4 t=tilesOf[k];
  uint32_t tt[4]={t.x,t.y,t.z,t.w};
  for (uint32_t j=0;j<c;j++)
  {
    pairKey[o+j]=tt[j];
    pairVal[o+j]=(uint32_t)k;
  }
}

__global__ void __launch_bounds__(256)
wb_segment_kernel(const uint32_t *__restrict__ pairKey,unsigned long long m,
                  uint32_t *__restrict__ tStart,uint32_t *__restrict__ tCount,
                  uint32_t *__restrict__ tileList,unsigned long long *__restrict__ nList)
// start and count of every tile's run in the sorted (tile,point) pairs + the list of non-empty tiles
{
  unsigned long long i=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  if (i>=m)
    return;
  const uint32_t t=pairKey[i];
  if (i==0 || pairKey[i-1]!=t)
  {
    tStart[t]=(uint32_t)i;
    unsigned long long lo=i+1,hi=m;  // first pair of the next tile: keys are sorted
    while (lo<hi)
    {
      unsigned long long mid=(lo+hi)>>1;
      if (pairKey[mid]<=t) lo=mid+1; else hi=mid;
    }
    tCount[t]=(uint32_t)(lo-i);
    tileList[atomicAdd(nList,1ull)]=(uint32_t)t;
  }
}

// ============================================================================ K7: tile scan
// scanCylinder (scan.cpp:31-140).  The 3x3 Gauss-Jordan below restates matrix.cpp's rowop /
// findpivot / gausselim; the kernel itself follows.

struct WbMat3 { double a[3][3],b[3]; };

__device__ void wb_swap_rows(WbMat3 &m,int r0,int r1)
{
  for (int i=0;i<3;i++)
  {
    double t=m.a[r0][i]; m.a[r0][i]=m.a[r1][i]; m.a[r1][i]=t;
  }
  double t=m.b[r0]; m.b[r0]=m.b[r1]; m.b[r1]=t;
}

__device__ double wb_pairwise_small(const double *a,int n)
// pairwisesum for n <= 3 (the squares vector of findpivot)
{
  if (n==1) return __dadd_rn(0.0,a[0]);
  if (n==2) return __dadd_rn(0.0,__dadd_rn(a[0],a[1]));
  return __dadd_rn(__dadd_rn(0.0,a[2]),__dadd_rn(a[0],a[1]));
}

__device__ void wb_rowop(WbMat3 &m,int row0,int row1,int piv)
// matrix::rowop, matrix.cpp:262-349
{
  int flags=0,pivot;
  double slope=0,minslope=INFINITY,detfactor;
  if (piv>=0 && m.a[row0][piv]==0 && m.a[row1][piv]==0)
    piv=-1;
  pivot=piv;
  if (piv>=0 && m.a[row0][piv]==0)
    flags=9;
  for (int i=0;piv<0 && i<3;i++)
    if (m.a[row0][i]!=0 || m.a[row1][i]!=0)
    {
      if (fabs(m.a[row0][i])>fabs(m.a[row1][i]) || row0>=row1)
      {
        slope=fabs(__ddiv_rn(m.a[row1][i],m.a[row0][i]));
        flags&=~8;
      }
      else
      {
        slope=fabs(__ddiv_rn(m.a[row0][i],m.a[row1][i]));
        flags|=8;
      }
      if (slope<minslope)
      {
        minslope=slope;
        flags=(flags>>3)*9;
        pivot=i;
      }
    }
  flags&=1;
  if (flags)
    wb_swap_rows(m,row0,row1);
  detfactor=pivot<0?0:m.a[row0][pivot];
  if (detfactor!=0 && detfactor!=1)
  {
    for (int i=0;i<3;i++)
      m.a[row0][i]=__ddiv_rn(m.a[row0][i],detfactor);
    m.b[row0]=__ddiv_rn(m.b[row0],detfactor);
  }
  if (pivot>=0)
    slope=m.a[row1][pivot];
  if (slope!=0 && row0!=row1)
  {
    for (int i=0;i<3;i++)
      m.a[row1][i]=__dsub_rn(m.a[row1][i],__dmul_rn(m.a[row0][i],slope));
    m.b[row1]=__dsub_rn(m.b[row1],__dmul_rn(m.b[row0],slope));
  }
}

__device__ void wb_findpivot(WbMat3 &m,int row,int column)
// matrix::findpivot, matrix.cpp:382-422
{
  int pivotrow=-1;
  double maxratio=0;
  for (;pivotrow<row && column<3;column++)
    for (int i=row;i<3;i++)
    {
      double sq[4]={0,0,0,0};
      for (int j=column+1;j<3;j++)
        sq[j-column-1]=__dmul_rn(m.a[i][j],m.a[i][j]);
      double ratio=__ddiv_rn(__dmul_rn(m.a[i][column],m.a[i][column]),wb_pairwise_small(sq,3-column));
      if (ratio>maxratio)
      {
        pivotrow=i;
        maxratio=ratio;
      }
    }
  if (pivotrow>row)
    wb_swap_rows(m,pivotrow,row);
}

__device__ void wb_gausselim(WbMat3 &m)
// matrix::gausselim, matrix.cpp:358-380
{
  for (int i=0;i<3;i++)
  {
    wb_findpivot(m,i,i);
    for (int j=0;j<3;j++)
      wb_rowop(m,i,j,i);
  }
  for (int i=2;i>=0;i--)
    for (int j=0;j<i;j++)
      wb_rowop(m,i,j,i);
}

#define WB_SCAN_WARPS 4
#define WB_SCAN_BUNDLE 32             // most tiles per warp

static inline int wb_scan_bundle(unsigned long long nTiles,unsigned long long nPairs)
// Tiles per warp.  Bundling pays for the plane fit (one lane per tile instead of 32 lanes on one tile), which matters
// when tiles are small (aerial: 10-60 points); a terrestrial scan has few, large tiles (C4: 69 k tiles of ~1 000
// points), where a warp that walks 32 of them one after the other leaves the GPU a few thousand warps to run:
// 162 ms against 24 ms with one tile per warp.  So: as many as keep about 64 points per bundled tile-batch and at
// least ~10 k warps in flight, a power of two, at most 32.
{
  if (const char *e=getenv("WB_SCAN_BUNDLE"))          // tests: small scenes would always get 1
  {
    const int v=atoi(e);
    if (v>=1 && v<=WB_SCAN_BUNDLE)
      return v;
  }
  if (!nTiles)
    return 1;
  const unsigned long long avg=nPairs/nTiles+1;
  unsigned long long b=2048/avg;
  b=b<nTiles/9472?b:nTiles/9472;
  int r=1;
  while (r*2<=WB_SCAN_BUNDLE && (unsigned long long)r*2<=b)
    r*=2;
  return r;
}

__global__ void __launch_bounds__(WB_SCAN_WARPS*32)
wb_scan_kernel(const uint32_t *__restrict__ tileList,uint32_t nList,
               const uint32_t *__restrict__ tStart,const uint32_t *__restrict__ tCount,
               const uint32_t *__restrict__ pairVal,
               const double *__restrict__ sx,const double *__restrict__ sy,const double *__restrict__ sz,
               WbSnake snake,double minHyp,int bundle,
               int *__restrict__ tNPoints,uint8_t *__restrict__ tTree,double *__restrict__ tDensity,
               double *__restrict__ tHyp,double *__restrict__ tHeight)
// scanCylinder (scan.cpp:31-140): ONE WARP PER BUNDLE OF `bundle` (1..32, wb_scan_bundle) NON-EMPTY TILES.
//   sums    tile after tile, the warp takes the tile's points (canonical order) 32 at a time, one per lane.  The eight
//           normal-equation sums keep the reference's association (pairwisesum, manysum.cpp:120-154 = perfect binary
//           trees over aligned power-of-two blocks, merged like a binary counter): inside a batch the tree is a
//           butterfly of warp shuffles (block sums of size 2^s are read off before step s for the tail of the last
//           batch); across batches lane j (j<8) carries quantity j's counter in shared memory.
//   plane   the 3x3 normal equations (matrix.cpp's Gauss-Jordan, a long scalar code full of divisions) are solved
//           with LANE = TILE, 32 tiles at once: one warp per tile solved the same system on all 32 lanes and spent
//           most of the kernel's issue slots there (aerial tiles hold 10-60 points: one or two batches).
//   layers  tile after tile again: lowest untilted point, bottom layer, seven-sector histogram.
{
  __shared__ double lvAll[WB_SCAN_WARPS][8][28];
  __shared__ double totAll[WB_SCAN_WARPS][WB_SCAN_BUNDLE][8];
  const int lane=threadIdx.x&31,wi=threadIdx.x>>5;
  const uint32_t first=(blockIdx.x*WB_SCAN_WARPS+wi)*(uint32_t)bundle;
  if (first>=nList)
    return;
  const int nT=(int)min((uint32_t)bundle,nList-first);
  double (*lv)[28]=lvAll[wi];
  // ---- lane = tile: where its points are, where its centre is
  uint32_t myT=0,myStart=0,myCnt=0;
  double myCx=0,myCy=0;
  if (lane<nT)
  {
    myT=tileList[first+lane];
    myStart=tStart[myT];
    myCnt=tCount[myT];
    if (myCnt>=(1u<<27))
      myCnt=(1u<<27)-1;               // 28 counter levels; unreachable below the 2^32-point limit
    int ex,ey;
    wb_to_flowsnake((int)myT+snake.lo,ex,ey);
    wb_tile_center(ex,ey,snake,myCx,myCy);
  }
  // ---- phase 1: pairwise sums of x*x, y*x, y*y, x, y, x*z, y*z, z  (1*1 sums to cnt exactly)
  for (int ti=0;ti<nT;ti++)
  {
    const uint32_t start=__shfl_sync(WB_FULL,myStart,ti),cnt=__shfl_sync(WB_FULL,myCnt,ti);
    const double ccx=__shfl_sync(WB_FULL,myCx,ti),ccy=__shfl_sync(WB_FULL,myCy,ti);
    const uint32_t nBatch=(cnt+31)>>5;
    for (uint32_t b=0;b<nBatch;b++)
    {
      const uint32_t i=b*32+lane;
      const bool valid=i<cnt;
      double v[8]={0,0,0,0,0,0,0,0};
      if (valid)
      {
        const uint32_t k=pairVal[start+i];
        const double x=__dsub_rn(sx[k],ccx),y=__dsub_rn(sy[k],ccy),z=sz[k];
        v[0]=__dmul_rn(x,x);
        v[1]=__dmul_rn(y,x);
        v[2]=__dmul_rn(y,y);
        v[3]=x;
        v[4]=y;
        v[5]=__dmul_rn(x,z);
        v[6]=__dmul_rn(y,z);
        v[7]=z;
      }
      const uint32_t r=min(32u,cnt-b*32);
      #pragma unroll
      for (int s=0;s<5;s++)
      {
        if (r<32 && ((r>>s)&1))
        {
          // the block of size 2^s of the tail starts where the higher bits of r end
          const int pos=(int)(r&~((2u<<s)-1));
          double mine=0;
          #pragma unroll
          for (int j=0;j<8;j++)
          {
            double e=__shfl_sync(WB_FULL,v[j],pos);
            if (lane==j)
              mine=e;
          }
          if (lane<8)
            lv[lane][s]=mine;
        }
        #pragma unroll
        for (int j=0;j<8;j++)
          v[j]=__dadd_rn(v[j],__shfl_xor_sync(WB_FULL,v[j],1<<s));
      }
      if (r==32)
      {
        double mine=0;
        #pragma unroll
        for (int j=0;j<8;j++)
          if (lane==j)
            mine=v[j];
        if (lane<8)
        {
          int l=5;
          uint32_t m=b;
          while (m&1)
          {
            mine=__dadd_rn(lv[lane][l],mine);
            m>>=1;
            l++;
          }
          lv[lane][l]=mine;
        }
      }
      __syncwarp();
    }
    if (lane<8)
    {
      double sum=0;
      for (int l=0;l<28;l++)
        if ((cnt>>l)&1)
          sum=__dadd_rn(sum,lv[lane][l]);
      totAll[wi][ti][lane]=sum;
    }
    __syncwarp();
  }
  // ---- phase 2: 3x3 normal equations, lane = tile
  double mySl0=0,mySl1=0;
  if (lane<nT)
  {
    const double *tot=totAll[wi][lane];
    WbMat3 m;
    m.a[0][0]=tot[0];
    m.a[1][0]=m.a[0][1]=tot[1];
    m.a[1][1]=tot[2];
    m.a[2][0]=m.a[0][2]=tot[3];
    m.a[2][1]=m.a[1][2]=tot[4];
    m.a[2][2]=(double)myCnt;
    m.b[0]=tot[5];
    m.b[1]=tot[6];
    m.b[2]=tot[7];
    wb_gausselim(m);
    double sl0=m.a[0][0]==0?NAN:m.b[0],sl1=m.a[1][1]==0?NAN:m.b[1];
    const double len=wb_hypot(sl0,sl1);
    if (len>1)
    {
      sl0=__ddiv_rn(sl0,len);
      sl1=__ddiv_rn(sl1,len);
    }
    if (isnan(sl0) || isnan(sl1))
      sl0=sl1=0;
    mySl0=sl0;
    mySl1=sl1;
  }
  __syncwarp();
  for (int ti=0;ti<nT;ti++)
  {
    const uint32_t t=__shfl_sync(WB_FULL,myT,ti),start=__shfl_sync(WB_FULL,myStart,ti),cnt=__shfl_sync(WB_FULL,myCnt,ti);
    const double ccx=__shfl_sync(WB_FULL,myCx,ti),ccy=__shfl_sync(WB_FULL,myCy,ti);
    const double sl0=__shfl_sync(WB_FULL,mySl0,ti),sl1=__shfl_sync(WB_FULL,mySl1,ti);
    const uint32_t nBatch=(cnt+31)>>5;
    // ---- phase 3: bottom = lowest untilted z, f = its first position, top; then
    //      bottom2 = lowest untilted z BEFORE position f (scan.cpp:78-87: the update only fires on a
    //      strict new minimum), = bottom if there is none
    double bestv=INFINITY,top=-INFINITY;
    uint32_t besti=0xffffffffu;
    double x0=0,y0=0,zu0=INFINITY;                  // the first batch stays in registers (most tiles have no other)
    for (uint32_t b=0;b<nBatch;b++)
    {
      const uint32_t i=b*32+lane;
      if (i<cnt)
      {
        const uint32_t k=pairVal[start+i];
        const double x=__dsub_rn(sx[k],ccx),y=__dsub_rn(sy[k],ccy);
        const double zu=__dsub_rn(sz[k],__dadd_rn(__dmul_rn(sl1,y),__dmul_rn(sl0,x)));   // dot(): a.y*b.y+a.x*b.x
        if (b==0)
        {
          x0=x;
          y0=y;
          zu0=zu;
        }
        if (zu<bestv)
        {
          bestv=zu;
          besti=i;
        }
        if (zu>top)
          top=zu;
      }
    }
    #pragma unroll
    for (int o=16;o;o>>=1)
    {
      const double ov=__shfl_xor_sync(WB_FULL,bestv,o);
      const uint32_t oi=__shfl_xor_sync(WB_FULL,besti,o);
      if (ov<bestv || (ov==bestv && oi<besti))
      {
        bestv=ov;
        besti=oi;
      }
      top=fmax(top,__shfl_xor_sync(WB_FULL,top,o));
    }
    const double bottom=bestv;
    double bottom2=INFINITY;
    if ((uint32_t)lane<besti && (uint32_t)lane<cnt)
      bottom2=zu0;
    for (uint32_t b=1;b*32<besti && b<nBatch;b++)
    {
      const uint32_t i=b*32+lane;
      if (i<besti && i<cnt)
      {
        const uint32_t k=pairVal[start+i];
        const double x=__dsub_rn(sx[k],ccx),y=__dsub_rn(sy[k],ccy);
        const double zu=__dsub_rn(sz[k],__dadd_rn(__dmul_rn(sl1,y),__dmul_rn(sl0,x)));
        bottom2=fmin(bottom2,zu);
      }
    }
    #pragma unroll
    for (int o=16;o;o>>=1)
      bottom2=fmin(bottom2,__shfl_xor_sync(WB_FULL,bottom2,o));
    if (isinf(bottom2))
      bottom2=bottom;
    // ---- phase 4: seven-sector histogram of the bottom layer
    int histo[7]={0,0,0,0,0,0,0};
    uint32_t nBottom=0;
    const double cut=__dadd_rn(bottom2,__dmul_rn(2.0,snake.radius));
    const double rin=__ddiv_rn(snake.radius,WB_SQRT7);
    for (uint32_t b=0;b<nBatch;b++)
    {
      const uint32_t i=b*32+lane;
      int sector=-1;
      if (i<cnt)
      {
        double x=x0,y=y0,zu=zu0;
        if (b)
        {
          const uint32_t k=pairVal[start+i];
          x=__dsub_rn(sx[k],ccx);
          y=__dsub_rn(sy[k],ccy);
          zu=__dsub_rn(sz[k],__dadd_rn(__dmul_rn(sl1,y),__dmul_rn(sl0,x)));
        }
        if (zu<cut)
        {
          sector=(int)wb_lrint(__ddiv_rn(__dmul_rn(atan2(y,x),3.0),WB_PI));
          if (sector<0)
            sector+=6;
          sector=(sector%6)+1;
          if (wb_hypot(x,y)<rin)
            sector=0;
        }
      }
      #pragma unroll
      for (int j=0;j<7;j++)
      {
        const int c=__popc(__ballot_sync(WB_FULL,sector==j));
        histo[j]+=c;
        nBottom+=c;
      }
    }
    if (lane==0)
    {
      double density=0;
      for (int j=0;j<7;j++)
        density=__dadd_rn(density,(double)(histo[j]*histo[j]));
      int tree=0;
      if (cnt>nBottom && density<7)
        tree=1;
      density=__ddiv_rn(__ddiv_rn(__dmul_rn(sqrt(density),WB_SQRT7),__dmul_rn(snake.radius,snake.radius)),WB_PI);
      if (cnt>nBottom && density<0.5)
        tree=1;
      if (__dsub_rn(top,bottom)>1.5)
        tree=1;
      tNPoints[t]=(int)cnt;
      tTree[t]=(uint8_t)tree;
      tDensity[t]=density;
      tHyp[t]=sqrt(__dadd_rn(__ddiv_rn(1.0,density),__dmul_rn(minHyp,minHyp)));
      tHeight[t]=__dsub_rn(top,bottom);
    }
  }
}

// ============================================================================ K8: postscan
// postscanCylinder (scan.cpp:142-179).  Reads only nPoints/treeFlags of other tiles, writes only
// this tile's hyperboloidSize: no ordering hazard.  A tile outside the table has nPoints == 0.
// The six ray walks step through Eisenstein addresses, so the populated tiles are first laid
// out as a dense byte grid over (ex,ey) (0 empty, 1 populated, 2 populated tree tile): a step
// is then one byte load instead of an inverse flowsnake numbering.

// The kernels run over the LIST of non-empty tiles (wb_segment_kernel), not over the dense table, which has 7^i
// entries (2.8e8 for the 1 B-point scene); [xlo,xhi) restricts them to the tiles a rank of a sharded run owns.

__device__ __forceinline__ bool wb_tile_owned(uint32_t t,const WbSnake &snake,double xlo,double xhi,int &ex,int &ey)
{
  wb_to_flowsnake((int)t+snake.lo,ex,ey);
  double cx,cy;
  wb_tile_center(ex,ey,snake,cx,cy);
  return cx>=xlo && cx<xhi;
}

__global__ void __launch_bounds__(256)
wb_tile_extent_list_kernel(const uint32_t *__restrict__ tileList,uint32_t nList,WbSnake snake,double xlo,double xhi,
                           int *__restrict__ ext)
// ext = {min ex, min ey, max ex, max ey} over the listed tiles whose centre x lies in [xlo,xhi)
{
  uint32_t i=blockIdx.x*blockDim.x+threadIdx.x;
  int x0=INT_MAX,y0=INT_MAX,x1=INT_MIN,y1=INT_MIN;
  if (i<nList)
  {
    int ex,ey;
    if (wb_tile_owned(tileList[i],snake,xlo,xhi,ex,ey))
    {
      x0=x1=ex;
      y0=y1=ey;
    }
  }
  #pragma unroll
  for (int o=16;o;o>>=1)
  {
    x0=min(x0,__shfl_xor_sync(WB_FULL,x0,o));
    y0=min(y0,__shfl_xor_sync(WB_FULL,y0,o));
    x1=max(x1,__shfl_xor_sync(WB_FULL,x1,o));
    y1=max(y1,__shfl_xor_sync(WB_FULL,y1,o));
  }
  if ((threadIdx.x&31)==0 && x0!=INT_MAX)
  {
    atomicMin(&ext[0],x0);
    atomicMin(&ext[1],y0);
    atomicMax(&ext[2],x1);
    atomicMax(&ext[3],y1);
  }
}

__global__ void __launch_bounds__(256)
wb_tile_grid_list_kernel(const uint32_t *__restrict__ tileList,uint32_t nList,const uint8_t *__restrict__ tTree,
                         WbSnake snake,double xlo,double xhi,const int *__restrict__ ext,uint8_t *__restrict__ grid)
{
  uint32_t i=blockIdx.x*blockDim.x+threadIdx.x;
  if (i>=nList)
    return;
  const uint32_t t=tileList[i];
  int ex,ey;
  if (!wb_tile_owned(t,snake,xlo,xhi,ex,ey))
    return;
  const long long W=(long long)ext[2]-ext[0]+1;
  grid[(long long)(ey-ext[1])*W+(ex-ext[0])]=(uint8_t)(1+(tTree[t]&1));
}

__global__ void __launch_bounds__(128)
wb_postscan_list_kernel(const uint32_t *__restrict__ tileList,uint32_t nList,const uint8_t *__restrict__ tTree,
                        WbSnake snake,const int *__restrict__ ext,const uint8_t *__restrict__ grid,double *__restrict__ tHyp)
// postscanCylinder (scan.cpp:142-179), one thread per listed tile; see wb_postscan_kernel
{
  uint32_t i=blockIdx.x*blockDim.x+threadIdx.x;
  if (i>=nList)
    return;
  const uint32_t t=tileList[i];
  const int rx[6]={1,1,0,-1,-1,0},ry[6]={0,1,1,0,-1,-1};       // root1, eisenstein.cpp:51
  int ex,ey,count=0;
  wb_to_flowsnake((int)t+snake.lo,ex,ey);
  if (tTree[t]&1)
  {
    const int x0=ext[0],y0=ext[1],x1=ext[2],y1=ext[3];
    const long long W=(long long)x1-x0+1;
    int k=1,ringcount,nontree;
    do
    {
      ringcount=nontree=0;
      #pragma unroll
      for (int j=0;j<6;j++)
      {
        int nx=ex+rx[j]*k,ny=ey+ry[j]*k;
        if (nx<x0 || nx>x1 || ny<y0 || ny>y1)
          continue;
        uint8_t c=grid[(long long)(ny-y0)*W+(nx-x0)];
        if (c)
        {
          ringcount++;
          if (c==2)
            count++;
          else
            nontree++;
        }
      }
      ++k;
    } while (ringcount && !nontree);
  }
  double h=tHyp[t],c=__ddiv_rn(__dmul_rn((double)count,snake.spacing),6.0);
  tHyp[t]=sqrt(__dadd_rn(__dmul_rn(h,h),__dmul_rn(c,c)));
}

__global__ void __launch_bounds__(256)
wb_max_hyp_list_kernel(const uint32_t *__restrict__ tileList,uint32_t nList,const double *__restrict__ tHyp,
                       WbSnake snake,double xlo,double xhi,unsigned long long *__restrict__ out)
// largest hyperboloidSize over the listed tiles with centre x in [xlo,xhi) (positive: bit pattern orders like the value)
{
  unsigned long long m=0;
  for (uint32_t i=blockIdx.x*blockDim.x+threadIdx.x;i<nList;i+=gridDim.x*blockDim.x)
  {
    const uint32_t t=tileList[i];
    int ex,ey;
    if (wb_tile_owned(t,snake,xlo,xhi,ex,ey))
    {
      double h=tHyp[t];
      if (h>0 && h<INFINITY)
        m=max(m,(unsigned long long)__double_as_longlong(h));
    }
  }
  #pragma unroll
  for (int o=16;o;o>>=1)
    m=max(m,__shfl_xor_sync(WB_FULL,m,o));
  if ((threadIdx.x&31)==0 && m)
    atomicMax(out,m);
}

// ============================================================================ K9: classify
// classifyCylinder (classify.cpp:96-173) as a pure per-point function: label 1 iff the bearings
// dir(P,Q) of every cloud point Q inside P's downward hyperboloid (dist_xy != 0) leave no
// circular gap >= 144 degrees (surround(), classify.cpp:67-94), else 2.
//
// One warp owns one chunk of 32 consecutive points of the canonical order (spatial neighbours)
// as its 32 QUERY points.  The warp walks the bucket hierarchy once for all of them.
//   * lane = query:  per visited node, can my hyperboloid reach it (lowest point at nearest xy),
//                    and could it still tell me anything (see sectors below)?
//   * lane = point:  per visited chunk, the 32 chunk points sit in registers, the interested
//                    queries are broadcast one at a time from shared memory, every lane tests
//                    its point, and the outcome is combined with warp reductions.
// Bearings are first only binned into 64 sectors of 5.625 degrees (a few multiplies against
// tan(k*5.625), exact atan2i only within 1e-7 of a sector edge).  144 degrees = 25.6 sectors, so
//   longest empty run <= 23 sectors  =>  every gap < 25 sectors = 140.6 degrees: surrounded;
//   longest empty run >= 26 sectors  =>  a gap  > 26 sectors = 146.2 degrees: not surrounded;
// and a node whose angular extent lies inside already-occupied sectors cannot change the
// occupancy and is skipped.  Only queries whose longest empty run is 24 or 25 sectors need exact
// bearings: a second, much narrower walk computes atan2i for the points of the (at most four)
// sectors that bound such runs and measures the gap exactly.

#ifndef WB_CL_WARPS
#define WB_CL_WARPS 1                    // one warp per CTA: a slow warp strands nothing (measured 8,4,2,1 warps: 1189,1115,1079,974 ms)
#endif

// Single-precision shortcuts (both leave every decision to the exact double code when in doubt; the error
// bounds are pinned by tests/test_float_filters.py): sector binning of the bearings, and the
// (chunk, query) reach test at expansion.  0 = the double-precision originals.
#ifndef WB_CL_FSECTOR
#define WB_CL_FSECTOR 1
#endif
#ifndef WB_CL_FREACH
#define WB_CL_FREACH 1
#endif
#ifndef WB_CL_FSPAN
#define WB_CL_FSPAN 1
#endif
// Bulk re-filter of a chunk entry's waiting children after an occupancy change (B200, 100 M-point tile: 843 -> 832 ms
// alone; profiles/r2_classify_ab.md).
#ifndef WB_CL_REFILTER
#define WB_CL_REFILTER 1
#endif
// The per-query reach test at expansion for the children of EVERY level, not only for chunks, and with each query's
// still-open sectors taken into account — so that a child no single query can use is never pushed (node pops per
// warp 555 -> 273 on the bench scene, 843 -> 804 ms alone).
#ifndef WB_CL_XWANTS
#define WB_CL_XWANTS 1
#endif
// ... up to this level of children only (0 = chunks, as without XWANTS): near the root nearly every child fails the
// group test already and the per-query loop costs more than the pops it saves when hyperboloids are small
#ifndef WB_CL_XWANTS_MAXLEVEL
#define WB_CL_XWANTS_MAXLEVEL 7
#endif
// The pair loop's "does any point of this chunk matter to query q" vote repeats what the ballot in front of the
// loop established (a query stays in qm only if some point's sector span meets its open sectors): 1 = drop the vote.
#ifndef WB_CL_NORELVOTE
#define WB_CL_NORELVOTE 1
#endif
// Per-query float copies as one float4 (one 16-byte shared load per asker) and the loop-invariant slack folded into
// the child's box before the asker loop.
#ifndef WB_CL_F4
#define WB_CL_F4 1
#endif
// The (child, query) reach tests of an expansion form a matrix: 32 children x the asking queries.  It is walked either
// with lanes = children and the askers broadcast one by one from shared memory (WB_CL_ACOST instructions each), or
// with lanes = queries and the children that passed the group test broadcast one by one by shuffles (WB_CL_TCOST
// each) — whichever is shorter for this node.  Near the root and for small hyperboloids two or three children pass
// the group test while all 32 queries ask; at chunk level it is the other way round.  0 = always lanes = children.
#ifndef WB_CL_TRANSPOSE
#define WB_CL_TRANSPOSE 1
#endif
#ifndef WB_CL_TCOST
#define WB_CL_TCOST 18
#endif
#ifndef WB_CL_ACOST
#define WB_CL_ACOST 14
#endif
#if WB_CL_TRANSPOSE && !(WB_CL_F4 && WB_CL_XWANTS)
#error "WB_CL_TRANSPOSE builds on WB_CL_F4 and WB_CL_XWANTS"
#endif
#if WB_CL_XWANTS && !WB_CL_FREACH
#error "WB_CL_XWANTS builds on the single-precision reach test (WB_CL_FREACH)"
#endif

// Pass 2 over COMPACTED pending queries.  On the bench scene 9.6 % of the queries need the exact walk but they sit
// in 65 % of the warps, a handful each; gathered (in canonical order, so still neighbours) into full warps the same
// walks are shared by 32 queries instead of ~5 (843 -> 793 ms alone; all three together 750 ms, same labels).
#ifndef WB_CL_COMPACT2
#define WB_CL_COMPACT2 1
#endif

// A popped INTERNAL node asks every live query again whether it reaches the node and still needs its sectors (double
// reach test + silhouette span, all lanes).  The expansion that pushed the node already knows which queries wanted
// it (the per-query float test of WB_CL_XWANTS): 1 = keep that set per stack entry and ask only "still live?" at the
// pop (stale sectors are caught by the next expansion's per-query test); 2 = also drop the queries whose open sectors
// miss the node's span as seen from the group (one span for the warp instead of one per query).  0 = ask again.
#ifndef WB_CL_POPWANTS
#define WB_CL_POPWANTS (WB_CL_XWANTS?1:0)
#endif
#if WB_CL_POPWANTS && !WB_CL_XWANTS
#error "WB_CL_POPWANTS builds on WB_CL_XWANTS"
#endif

#ifndef WB_EMU_COUNT
#define WB_EMU_COUNT(slot)              // loop-trip counters of the SIMT emulator (tests/simt); nothing on the GPU
#endif

struct WbClassifyWarp
{
  double qx[32],qy[32],qcz[32],qpor2[32];
#if WB_CL_FREACH
#if WB_CL_XWANTS
  unsigned long long openq[32];         // each query's sectors that can still matter (0 once it is decided)
#endif
#if WB_CL_F4
  float4 fq[32];                        // the same queries relative to the warp's origin: x, y, vertex height, 2*por
#else
  float fx[32],fy[32],fh[32],f2p[32];   // the same queries relative to the warp's origin: xy, vertex height, 2*por
#endif
  double org[3];                        // that origin (the warp's first query)
  float fgh,fzq;                        // bounds of |fx|,|fy| and of |fh| over the warp's queries
#endif
  uint32_t keys[8][32];       // per stack entry: (squared distance | child) of the children still to visit
#if WB_CL_POPWANTS
  uint32_t wantsLv[8][32];    // ... and the queries that wanted each child when it was pushed
#endif
  WbBound cb[32];             // bounds of the chunks of the open level-0 entry
  uint32_t wants[32];         // ... live queries that reach each chunk
  unsigned long long cm[32];  // ... and the sectors the chunk can occupy, seen from anywhere in the group
  uint32_t stBase[8];
  int stLevel[8];
};

__device__ __forceinline__ bool wb_reach(double gx0,double gx1,double gy0,double gy1,double cz,double por2,
                                         double s2,const WbBound &b)
// Could a hyperboloid with centre height cz, squared polar radius por2 and vertex xy anywhere in
// [gx0,gx1]x[gy0,gy1] contain a point of the box b?  Conservative (slack 1e-12 relative).
{
  double dx=fmax(0.0,fmax(b.xmin-gx1,gx0-b.xmax));
  double dy=fmax(0.0,fmax(b.ymin-gy1,gy0-b.ymax));
  double zl=cz-b.zmin;
  if (!(zl>0))
    return false;
  double d2=(dx*dx+dy*dy)*s2;
  return zl*zl*(1+1e-12)-d2*(1-1e-12)>=por2*(1-1e-12);
}

__device__ __forceinline__ int wb_sector64(double dx,double dy)
// Sector (0..63, 5.625 degrees each, counter-clockwise from +x) of the bearing whose binary angle
// is u = atan2i(dy,dx) & 0x7fffffff, i.e. u >> 25; -1 if the direction is within 1e-7 (relative)
// of a sector edge and the exact atan2i has to decide.
{
  const double T1=0.09849140335716425,T2=0.198912367379658,T3=0.3033466836073424,T4=0.41421356237309503,
               T5=0.5345111359507916,T6=0.6681786379192989,T7=0.8206787908286602;
  double ax=fabs(dx),ay=fabs(dy);
  bool sw=ay>ax;
  double lo=sw?ax:ay,hi=sw?ay:ax;
  double c=hi*T4;
  bool b1=lo>=c;
  double m=fabs(lo-c);
  c=hi*(b1?T6:T2);
  bool b2=lo>=c;
  m=fmin(m,fabs(lo-c));
  c=hi*(b1?(b2?T7:T5):(b2?T3:T1));
  bool b3=lo>=c;
  m=fmin(m,fabs(lo-c));
  m=fmin(m,fmin(lo,hi-lo));
  if (!(m>1e-7*hi))                           // ~30 units of 2^-31 turn: atan2i rounds to integers
    return -1;
  int sub=(b1?4:0)+(b2?2:0)+(b3?1:0);
  int s1=sw?15-sub:sub;
  if (dy>=0)
    return dx>=0?s1:31-s1;
  return dx<0?32+s1:63-s1;
}

__device__ __forceinline__ int wb_sector64f(double dxd,double dyd)
// wb_sector64 with the three tangent comparisons in single precision (half the instructions: a
// double select is two moves).  The inputs lose 2^-24 each in the conversion and every product
// another 2^-24, so a comparison lo >= hi*T is off by at most 2.5e-7*hi; anything within 4e-6*hi
// of a sector edge returns -1 and the exact atan2i decides, as before.  Quadrants come from the
// signs of the doubles.
{
  const float T1=0.09849140335716425f,T2=0.198912367379658f,T3=0.3033466836073424f,T4=0.41421356237309503f,
              T5=0.5345111359507916f,T6=0.6681786379192989f,T7=0.8206787908286602f;
  const float ax=fabsf((float)dxd),ay=fabsf((float)dyd);
  const bool sw=ay>ax;
  const float lo=sw?ax:ay,hi=sw?ay:ax;
  float c=hi*T4;
  const bool b1=lo>=c;
  float m=fabsf(lo-c);
  c=hi*(b1?T6:T2);
  const bool b2=lo>=c;
  m=fminf(m,fabsf(lo-c));
  c=hi*(b1?(b2?T7:T5):(b2?T3:T1));
  const bool b3=lo>=c;
  m=fminf(m,fabsf(lo-c));
  m=fminf(m,fminf(lo,hi-lo));
  if (!(m>4e-6f*hi) || !(hi>1e-30f) || !(hi<1e30f))     // near an edge, or outside the range where float keeps the ratio
    return -1;
  const int sub=(b1?4:0)+(b2?2:0)+(b3?1:0);
  const int s1=sw?15-sub:sub;
  if (dyd>=0)
    return dxd>=0?s1:31-s1;
  return dxd<0?32+s1:63-s1;
}

__device__ __forceinline__ int wb_sector64_exact(double dx,double dy,uint32_t &u)
{
  u=(uint32_t)wb_atan2i(dy,dx)&0x7fffffffu;
  return (int)(u>>25);
}

__device__ __forceinline__ unsigned long long wb_rotl64(unsigned long long x,int r)
{
  r&=63;
  return r?(x<<r)|(x>>(64-r)):x;
}

__device__ __forceinline__ float wb_fast_angle(float x,float y)
// Bearing of (x,y) in sector units [0,64], accurate to 0.04 sector (atan(q) ~ q(pi/4+0.273(1-q))).
// Only used to build conservative masks; every decision about a real bearing uses wb_sector64.
{
  float ax=fabsf(x),ay=fabsf(y);
  float mn=fminf(ax,ay),mx=fmaxf(ax,ay);
  float q=__fdividef(mn,mx);
  float a=q*(8.0f+2.781f*(1.0f-q));
  if (ay>ax) a=16.0f-a;
  if (x<0) a=32.0f-a;
  if (y<0) a=64.0f-a;
  return a;
}

__device__ __forceinline__ unsigned long long wb_span_mask(double dx0,double dx1,double dy0,double dy1)
// Sectors of all bearings from the origin to the rectangle [dx0,dx1]x[dy0,dy1]: the two
// silhouette corners, each widened by 0.06 sector for the approximate angle.
{
#if WB_CL_FSPAN
  // the corners are only ever used as floats: convert first, so that the selects move one register, not two
  // (the conversion keeps signs; a difference that underflows to 0 makes the rectangle touch the origin: wider mask)
  const float x0=(float)dx0,x1=(float)dx1,y0=(float)dy0,y1=(float)dy1;
  const bool L=x0>0,R=x1<0,B=y0>0,T=y1<0;
  if (!(L||R||B||T))
    return ~0ull;                                  // the origin is inside
  const float ax=B?x1:(T?x0:(L?x0:x1)),ay=L?y0:(R?y1:(B?y0:y1));
  const float bx=B?x0:(T?x1:(L?x0:x1)),by=L?y1:(R?y0:(B?y0:y1));
#else
  const bool L=dx0>0,R=dx1<0,B=dy0>0,T=dy1<0;
  if (!(L||R||B||T))
    return ~0ull;                                  // the origin is inside
  // clockwise-most corner a, counter-clockwise-most corner b
  float ax=(float)(B?dx1:(T?dx0:(L?dx0:dx1))),ay=(float)(L?dy0:(R?dy1:(B?dy0:dy1)));
  float bx=(float)(B?dx0:(T?dx1:(L?dx0:dx1))),by=(float)(L?dy1:(R?dy0:(B?dy0:dy1)));
#endif
  int sa=(int)floorf(wb_fast_angle(ax,ay)-0.06f),sb=(int)floorf(wb_fast_angle(bx,by)+0.06f);
  int len=((sb-sa)&63)+1;
  if (len>40)                                      // a rectangle spans < 180 degrees; anything else is a wrap artefact
    return ~0ull;
  return wb_rotl64((1ull<<len)-1,sa&63);
}

__device__ __forceinline__ unsigned long long wb_runs_ge(unsigned long long empty,int len)
// bit i set iff sectors i-len+1..i (circular) are all empty; len in {24,26}
{
  unsigned long long r=empty;
  r&=wb_rotl64(r,1);
  r&=wb_rotl64(r,2);
  r&=wb_rotl64(r,4);
  r&=wb_rotl64(r,8);                           // runs of 16
  return r&wb_rotl64(r,len-16);
}

__device__ __forceinline__ unsigned long long wb_long_runs(unsigned long long occ)
// Sectors that belong to an empty run of at least 24 sectors.  Only these can still decide the
// outcome: shorter empty runs stay short whatever else turns up.
{
  unsigned long long x=wb_runs_ge(~occ,24);     // ends of such runs
  x|=wb_rotl64(x,63);
  x|=wb_rotl64(x,62);
  x|=wb_rotl64(x,60);
  x|=wb_rotl64(x,56);                           // each end spread back over 16 sectors
  return x|wb_rotl64(x,56);                     // ... over 24
}

__device__ __forceinline__ bool wb_in_hyperboloid(double px,double py,double pcz,double ppor2,double s2,double maxSlope,
                                                  double qx,double qy,double qz,double &ddx,double &ddy,bool &margin)
// Hyperboloid::in (shape.cpp:127-135) for the hyperboloid of query P and cloud point Q, and
// dist_xy(P,Q) != 0 (classify.cpp:150).  ddx,ddy = Q-P (the vector dir() takes the bearing of).
{
  double zd=__dsub_rn(pcz,qz);                      // centre.z - pnt.z
  if (!(zd>0))
    return false;
  double dx=__dsub_rn(px,qx),dy=__dsub_rn(py,qy);
  ddx=-dx;
  ddy=-dy;
  double zz=zd*zd,hs2=(dx*dx+dy*dy)*s2;
  double diff=zz-hs2-ppor2,tol=1e-13*(zz+hs2+ppor2);
  bool in;
  if (diff>tol)
    in=true;
  else if (diff<-tol)
    in=false;
  else
  {
    // too close to call without the reference's exact expression
    double d=wb_hypot(dx,dy);
    double ds=__dmul_rn(d,maxSlope);
    double lhs=__dsub_rn(__dmul_rn(zd,zd),__dmul_rn(ds,ds));
    in=lhs>=ppor2;
    if (fabs(lhs-ppor2)<=1e-12*(zz+ppor2) && (dx!=0 || dy!=0))
      margin=true;
  }
  return in && (dx!=0 || dy!=0);
}

#ifndef WB_CL_MINBLOCKS
#define WB_CL_MINBLOCKS (32/WB_CL_WARPS)   // 64 registers, 32 resident warps per SM: measured best of 16..48 warps
#endif
template <int PASS>
__global__ void __launch_bounds__(WB_CL_WARPS*32,WB_CL_MINBLOCKS)
wb_classify_kernel(const double *__restrict__ sx,const double *__restrict__ sy,const double *__restrict__ sz,
                   unsigned long long n,uint32_t nChunks,
                   const WbBound *__restrict__ bounds,const uint32_t *__restrict__ levelOff,
                   const uint32_t *__restrict__ levelCnt,int nLevels,
                   const uint32_t *__restrict__ winner,const double *__restrict__ tHyp,
                   double maxSlope,double thickness,
                   const uint8_t *__restrict__ clsIn,const uint32_t *__restrict__ perm,
                   uint32_t ownFirst,uint32_t ownEnd,const uint8_t *__restrict__ forcedIn,
                   uint8_t *__restrict__ labelSorted,unsigned long long *__restrict__ counters,
                   uint32_t *__restrict__ wedgeBuf,uint8_t *__restrict__ chunkPending
#if WB_CL_COMPACT2
                   ,const uint32_t *__restrict__ pendingList,uint32_t nPendingQ   // pass 2: the queries to walk for
#endif
                   )
// PASS 1: sector walk, decides every query whose longest empty run is not 24 or 25 sectors and
//         leaves the bounding sectors of the others in wedgeBuf (chunkPending marks their chunks).
// PASS 2: exact walk for the pending queries only.
{
  __shared__ WbClassifyWarp wsh[WB_CL_WARPS];
  WbClassifyWarp &w=wsh[threadIdx.x>>5];
  const int lane=threadIdx.x&31;
  const uint32_t chunk=blockIdx.x*WB_CL_WARPS+(threadIdx.x>>5);
#if WB_CL_COMPACT2
  unsigned long long me;
  bool have;
  if (PASS==2)
  {
    const unsigned long long slot=(unsigned long long)chunk*32+lane;
    if ((unsigned long long)chunk*32>=nPendingQ)
      return;
    have=slot<nPendingQ;
    me=have?pendingList[slot]:0;
  }
  else
  {
    if (chunk>=nChunks)
      return;
    me=(unsigned long long)chunk*32+lane;
    have=me<n;
  }
#else
  if (chunk>=nChunks)
    return;
  if (PASS==2 && !chunkPending[chunk])
    return;
  const unsigned long long me=(unsigned long long)chunk*32+lane;
  const bool have=me<n;
#endif
  const double s2=maxSlope*maxSlope;
  double px=0,py=0,pcz=-INFINITY,ppor2=INFINITY;
  bool done=true,untiled=false,foreign=false;
  if (have)
  {
    px=sx[me];
    py=sy[me];
    uint32_t src=perm[me];
    foreign=src<ownFirst || src>=ownEnd;          // halo point of another GPU: not ours to label ...
    if (foreign && forcedIn && forcedIn[src])
      foreign=false;                              // ... unless it holds the place of one of our records (same XYZ)
    uint32_t wt=winner[me];
    if (foreign)
      ;
    else if (wt!=0xffffffffu)
    {
      double r=tHyp[wt];
      double por=__dmul_rn(r,__dmul_rn(maxSlope,maxSlope));       // Hyperboloid ctor, shape.cpp:119-125
      ppor2=__dmul_rn(por,por);
      pcz=__dadd_rn(__dsub_rn(sz[me],thickness),por);
      done=false;
    }
    else
      untiled=true;
  }
  w.qx[lane]=px; w.qy[lane]=py; w.qcz[lane]=pcz; w.qpor2[lane]=ppor2;
#if WB_CL_FREACH
  // Single-precision copies for the (chunk, query) reach test at expansion: coordinates relative to the
  // first query of the warp, so that the conversion error is 2^-24 of a distance, not of an easting.
  {
    const double ox=__shfl_sync(WB_FULL,px,0),oy=__shfl_sync(WB_FULL,py,0),oz=__shfl_sync(WB_FULL,have?sz[me]:0.0,0);
    const bool q=have && !done;
    const double por=q?sqrt(ppor2):0.0;
    const float fx=q?(float)(px-ox):0.0f,fy=q?(float)(py-oy):0.0f,fh=q?(float)((pcz-por)-oz):0.0f;
#if WB_CL_F4
    w.fq[lane]=make_float4(fx,fy,fh,fminf((float)(2*por),1e30f));
#else
    w.fx[lane]=fx; w.fy[lane]=fy; w.fh[lane]=fh; w.f2p[lane]=fminf((float)(2*por),1e30f);
#endif
    const float gh=__uint_as_float(__reduce_max_sync(WB_FULL,__float_as_uint(fmaxf(fabsf(fx),fabsf(fy)))));
    const float zq=__uint_as_float(__reduce_max_sync(WB_FULL,__float_as_uint(fabsf(fh))));
    if (lane==0)
    {
      w.org[0]=ox; w.org[1]=oy; w.org[2]=oz;
      w.fgh=gh; w.fzq=zq;
    }
  }
#endif
  unsigned long long occ=0;                       // occupied sectors of my query
  unsigned long long open=~0ull;                  // sectors of empty runs >= 24 (the only ones that matter)
  bool changed=false;
  uint32_t statNodes=0,statChunks=0,statPairs=0,statNodes2=0,statChunks2=0,statPairs2=0;  // work counters (warp-uniform)
  bool margin=false,surrounded=false;
  // second-walk state: up to two empty runs of 24/25 sectors, bounded below by sector k1 (we need
  // the largest bearing in it) and above by k2 (the smallest bearing in it)
  uint32_t wedge=0xffffffffu;                     // k1a | k2a<<8 | k1b<<16 | k2b<<24, 0xff = none
  uint32_t maxLowA=0,minHighA=0xffffffffu,maxLowB=0,minHighB=0xffffffffu;
  unsigned long long wedgeMask=0;
  const int top=nLevels-1;
  constexpr int pass=PASS;
  if (PASS==2)
  {
    if (have && !done)
      wedge=wedgeBuf[me];
    for (int k=0;k<4;k++)
    {
      uint32_t sct=(wedge>>(8*k))&255;
      if (sct!=255)
        wedgeMask|=1ull<<sct;
    }
  }
  {
    bool live=pass==1?!done:wedgeMask!=0;
    uint32_t liveMask=__ballot_sync(WB_FULL,live);
    if (!liveMask)
      goto finish;
    // group envelope over the live queries: xy box, highest centre, smallest polar radius,
    // and the union of the sectors any of them still cares about
    double gx0,gx1,gy0,gy1,gcz,gpor2,gmx,gmy;
    unsigned long long needAny;
    uint32_t envMask=0;
    auto envelope=[&]()
    {
      gx0=live?px:INFINITY; gx1=live?px:-INFINITY;
      gy0=live?py:INFINITY; gy1=live?py:-INFINITY;
      gcz=live?pcz:-INFINITY; gpor2=live?ppor2:INFINITY;
      #pragma unroll
      for (int o=16;o;o>>=1)
      {
        gx0=fmin(gx0,__shfl_xor_sync(WB_FULL,gx0,o));
        gx1=fmax(gx1,__shfl_xor_sync(WB_FULL,gx1,o));
        gy0=fmin(gy0,__shfl_xor_sync(WB_FULL,gy0,o));
        gy1=fmax(gy1,__shfl_xor_sync(WB_FULL,gy1,o));
        gcz=fmax(gcz,__shfl_xor_sync(WB_FULL,gcz,o));
        gpor2=fmin(gpor2,__shfl_xor_sync(WB_FULL,gpor2,o));
      }
      gmx=0.5*(gx0+gx1); gmy=0.5*(gy0+gy1);
      envMask=liveMask;
    };
    auto needed=[&]()
    {
      unsigned long long mine=live?(pass==1?open:wedgeMask):0ull;
      uint32_t lo=__reduce_or_sync(WB_FULL,(uint32_t)mine),hi=__reduce_or_sync(WB_FULL,(uint32_t)(mine>>32));
      needAny=(unsigned long long)lo|((unsigned long long)hi<<32);
    };
    envelope();
    needed();
#if WB_CL_XWANTS
    w.openq[lane]=live?(pass==1?open:wedgeMask):0ull;
    __syncwarp();
#endif
    // lanes = children of a node: can any live query reach it, and does any still need its sectors?
    auto childTest=[&](const WbBound &cb,uint32_t &key,unsigned long long &cm)->bool
    {
      if (!wb_reach(gx0,gx1,gy0,gy1,gcz,gpor2,s2,cb))
        return false;
      double mx=0.5*(cb.xmin+cb.xmax)-gmx,my=0.5*(cb.ymin+cb.ymax)-gmy;
      float d2=(float)(mx*mx+my*my);
      key=(__float_as_uint(d2)&0xffffffe0u)|(uint32_t)lane;
      cm=wb_span_mask(cb.xmin-gx1,cb.xmax-gx0,cb.ymin-gy1,cb.ymax-gy0);   // valid for every query of the group
      return (cm&needAny)!=0;
    };
    // Push the children [base,base+32) of a node (level childLevel).  For chunks (childLevel 0)
    // each child also gets the set of live queries whose hyperboloid reaches it, found with
    // lanes = children and the queries broadcast one by one, so that popping a chunk is cheap.
    int sp=0;
    auto expand=[&](int childLevel,uint32_t base,uint32_t askers)
    {
      uint32_t c=base+lane,cc=levelCnt[childLevel];
      uint32_t key=0xffffffffu,wants=0;
      unsigned long long cm=0;
      bool ok=false;
      WbBound cb;
      if (c<cc)
      {
        cb=bounds[levelOff[childLevel]+c];
        ok=childTest(cb,key,cm);
      }
      if (!__any_sync(WB_FULL,ok))
        return;
      if (childLevel==0 || (WB_CL_XWANTS && childLevel<=WB_CL_XWANTS_MAXLEVEL))
      {
        uint32_t lm=askers;
#if WB_CL_FREACH
        // Conservative single-precision form of wb_reach(query, chunk box): with a = vertex height - lowest z,
        // (a+por)^2 - d^2 s^2 >= por^2  <=>  a (a + 2 por) >= d^2 s^2  (a >= 0).  Every rounding is covered: the
        // distances shrink by ed, the drop grows by ez (8x the worst conversion + subtraction error), the
        // products carry 4e-6 (their own roundings stay below 1e-6).  It only has to admit every pair the double test admits; the chunk's points are
        // tested exactly later.
        const double ox=w.org[0],oy=w.org[1],oz=w.org[2];
        const float x0=(float)(cb.xmin-ox),x1=(float)(cb.xmax-ox),y0=(float)(cb.ymin-oy),y1=(float)(cb.ymax-oy),
                    z0=(float)(cb.zmin-oz);
        const float ed=9.5367431640625e-7f*(fmaxf(fmaxf(fabsf(x0),fabsf(x1)),fmaxf(fabsf(y0),fabsf(y1)))+w.fgh);
        const float ez=9.5367431640625e-7f*(fabsf(z0)+w.fzq)+1e-6f;
#if WB_CL_F4
        // the slack goes into the box once: max(x0-qx,qx-x1)-ed = max((x0-ed)-qx,qx-(x1+ed)); a = qh-(z0-ez);
        // a(a+2por)*1.000002 >= d2*s2*0.999998  <=>  a(a+2por) >= d2*k2 with k2 rounded DOWN (admits a superset)
        const float bx0=x0-ed,bx1=x1+ed,by0=y0-ed,by1=y1+ed,bz0=z0-ez;
        const float k2=(float)s2*0.999995f;
#if WB_CL_TRANSPOSE
        const uint32_t okm=__ballot_sync(WB_FULL,ok);
        if (__popc(okm)*WB_CL_TCOST<__popc(askers)*WB_CL_ACOST)
        {
          // lanes = queries: the same test, the child's box (slack included) and sector span arriving by shuffle
          const float4 me=w.fq[lane];
          const unsigned long long myOpen=w.openq[lane];
          const bool asking=(askers>>lane)&1;
          uint32_t rem=okm;
          lm=0;
          while (rem)
          {
            const int c=__ffs(rem)-1;
            rem&=rem-1;
            WB_EMU_COUNT(childLevel==0?0:1);
            const float cx0=__shfl_sync(WB_FULL,bx0,c),cx1=__shfl_sync(WB_FULL,bx1,c);
            const float cy0=__shfl_sync(WB_FULL,by0,c),cy1=__shfl_sync(WB_FULL,by1,c),cz0=__shfl_sync(WB_FULL,bz0,c);
            const uint32_t clo=__shfl_sync(WB_FULL,(uint32_t)cm,c),chi=__shfl_sync(WB_FULL,(uint32_t)(cm>>32),c);
            const float dx=fmaxf(0.0f,fmaxf(cx0-me.x,me.x-cx1));
            const float dy=fmaxf(0.0f,fmaxf(cy0-me.y,me.y-cy1));
            const float a=me.z-cz0;
            const bool t=asking && a>=0.0f && a*(a+me.w)>=(dx*dx+dy*dy)*k2 &&
                         ((clo&(uint32_t)myOpen)|(chi&(uint32_t)(myOpen>>32)))!=0;
            const uint32_t m=__ballot_sync(WB_FULL,t);
            if (lane==c)
              wants=m;
          }
        }
#endif
        while (lm)
        {
          const int q=__ffs(lm)-1;
          lm&=lm-1;
          WB_EMU_COUNT(childLevel==0?0:1);
          const float4 fq=w.fq[q];
          const float dx=fmaxf(0.0f,fmaxf(bx0-fq.x,fq.x-bx1));
          const float dy=fmaxf(0.0f,fmaxf(by0-fq.y,fq.y-by1));
          const float a=fq.z-bz0;
#if WB_CL_XWANTS
          if (ok && a>=0.0f && a*(a+fq.w)>=(dx*dx+dy*dy)*k2 && (cm&w.openq[q])!=0)
#else
          if (ok && a>=0.0f && a*(a+fq.w)>=(dx*dx+dy*dy)*k2)
#endif
            wants|=1u<<q;
        }
#else
        const float fs2=(float)s2*0.999998f;
        while (lm)
        {
          const int q=__ffs(lm)-1;
          lm&=lm-1;
          WB_EMU_COUNT(childLevel==0?0:1);
          const float qx=w.fx[q],qy=w.fy[q],qh=w.fh[q],q2p=w.f2p[q];
          const float dx=fmaxf(0.0f,fmaxf(x0-qx,qx-x1)-ed);
          const float dy=fmaxf(0.0f,fmaxf(y0-qy,qy-y1)-ed);
          const float a=(qh-z0)+ez;
#if WB_CL_XWANTS
          if (ok && a>=0.0f && a*(a+q2p)*1.000002f>=(dx*dx+dy*dy)*fs2 && (cm&w.openq[q])!=0)
#else
          if (ok && a>=0.0f && a*(a+q2p)*1.000002f>=(dx*dx+dy*dy)*fs2)
#endif
            wants|=1u<<q;
        }
#endif
#else
        while (lm)
        {
          const int q=__ffs(lm)-1;
          lm&=lm-1;
          const double qx=w.qx[q],qy=w.qy[q],qcz=w.qcz[q],qpor2=w.qpor2[q];
          if (ok && wb_reach(qx,qx,qy,qy,qcz,qpor2,s2,cb))
            wants|=1u<<q;
        }
#endif
        ok=ok && wants!=0;
        if (!__any_sync(WB_FULL,ok))
          return;
        if (childLevel==0)
        {
          w.wants[lane]=wants;
          w.cm[lane]=cm;
          w.cb[lane]=cb;
        }
      }
      w.keys[sp][lane]=ok?key:0xffffffffu;
#if WB_CL_POPWANTS
      w.wantsLv[sp][lane]=(childLevel==0 || (WB_CL_XWANTS && childLevel<=WB_CL_XWANTS_MAXLEVEL))?wants:askers;
#endif
      if (lane==0)
      {
        w.stLevel[sp]=childLevel;
        w.stBase[sp]=base;
      }
      sp++;
      __syncwarp();
    };
    expand(top,0,liveMask);
    while (sp>0)
    {
      // pop the nearest remaining child of the top entry
      uint32_t best=__reduce_min_sync(WB_FULL,w.keys[sp-1][lane]);
      if (best==0xffffffffu)
      {
        sp--;
        continue;
      }
      const int bit=best&31;
      const int level=w.stLevel[sp-1];
      const uint32_t node=w.stBase[sp-1]+bit;
      if (lane==bit)
        w.keys[sp-1][lane]=0xffffffffu;
      __syncwarp();
      if (pass==1) statNodes++; else statNodes2++;
      // lane = query: does this node (still) matter to me?
      bool want=false;
      if (level>0)
      {
#if WB_CL_POPWANTS
        uint32_t askers=w.wantsLv[sp-1][bit]&liveMask;
#if WB_CL_POPWANTS==2
        if (askers)
        {
          const WbBound nb=bounds[levelOff[level]+node];
          const unsigned long long bs=wb_span_mask(nb.xmin-gx1,nb.xmax-gx0,nb.ymin-gy1,nb.ymax-gy0);
          askers&=__ballot_sync(WB_FULL,(bs&(pass==1?open:wedgeMask))!=0);
        }
#endif
#else
        WbBound nb=bounds[levelOff[level]+node];
        if (live && wb_reach(px,px,py,py,pcz,ppor2,s2,nb))
        {
          unsigned long long bs=wb_span_mask(nb.xmin-px,nb.xmax-px,nb.ymin-py,nb.ymax-py);
          want=pass==1?(bs&open)!=0:(bs&wedgeMask)!=0;
        }
        const uint32_t askers=__ballot_sync(WB_FULL,want);
#endif
        if (askers)
          expand(level-1,node*32,askers);
        continue;
      }
      {
        const uint32_t wq=w.wants[bit];
        const unsigned long long cmv=w.cm[bit];
        want=live && ((wq>>lane)&1) && (cmv&(pass==1?open:wedgeMask))!=0;
      }
      if (!__any_sync(WB_FULL,want))
        continue;
      {
        // the chunk's bearings as seen from my own query: tight silhouette span
        const WbBound nb=w.cb[bit];
        if (want)
          want=(wb_span_mask(nb.xmin-px,nb.xmax-px,nb.ymin-py,nb.ymax-py)&(pass==1?open:wedgeMask))!=0;
      }
      uint32_t qm=__ballot_sync(WB_FULL,want);
      if (!qm)
        continue;
      // ---- a chunk: lane = point
      const unsigned long long j=(unsigned long long)node*32+lane;
      const bool okp=j<n;
      const double cxp=okp?sx[j]:0.0,cyp=okp?sy[j]:0.0,czp=okp?sz[j]:INFINITY;
      if (pass==1) { statChunks++; statPairs+=__popc(qm); } else { statChunks2++; statPairs2+=__popc(qm); }
      // Sectors each chunk point can occupy as seen from ANY query of the group (bearing from the
      // group centre, widened by the group radius).  A query for which none of them is still of
      // interest skips the chunk, and within the chunk only the points that can land in one of
      // the query's interesting sectors are tested at all.
      unsigned long long pmask=0;
      if (okp)
        pmask=wb_span_mask(cxp-gx1,cxp-gx0,cyp-gy1,cyp-gy0);
      {
        uint32_t lo=__reduce_or_sync(WB_FULL,(uint32_t)pmask),hi=__reduce_or_sync(WB_FULL,(uint32_t)(pmask>>32));
        unsigned long long cme=(unsigned long long)lo|((unsigned long long)hi<<32);
        qm&=__ballot_sync(WB_FULL,(cme&(pass==1?open:wedgeMask))!=0);
      }
      while (qm)
      {
        const int q=__ffs(qm)-1;
        qm&=qm-1;
#if WB_CL_XWANTS
        const unsigned long long oq=w.openq[q];      // = query q's open (pass 2: wedge) sectors; refreshed after every chunk
#else
        const unsigned long long mineq=pass==1?open:wedgeMask;
        const unsigned long long oq=((unsigned long long)__shfl_sync(WB_FULL,(uint32_t)(mineq>>32),q)<<32)|
                                    __shfl_sync(WB_FULL,(uint32_t)mineq,q);
#endif
        const bool rel=(pmask&oq)!=0;
#if !WB_CL_NORELVOTE
        if (!__any_sync(WB_FULL,rel))
          continue;
#endif
        const double qx=w.qx[q],qy=w.qy[q],qcz=w.qcz[q],qpor2=w.qpor2[q];
        double ddx=0,ddy=0;
        bool in=rel && wb_in_hyperboloid(qx,qy,qcz,qpor2,s2,maxSlope,cxp,cyp,czp,ddx,ddy,margin);
        if (!__any_sync(WB_FULL,in))
          continue;
        if (pass==1)
        {
          int s=-2;
          if (in)
          {
            s=WB_CL_FSECTOR?wb_sector64f(ddx,ddy):wb_sector64(ddx,ddy);
            if (s<0)
            {
              uint32_t u;
              s=wb_sector64_exact(ddx,ddy,u);
            }
          }
          uint32_t lo=__reduce_or_sync(WB_FULL,(in && s<32)?1u<<s:0u);
          uint32_t hi=__reduce_or_sync(WB_FULL,(in && s>=32)?1u<<(s-32):0u);
          if (lane==q)
          {
            unsigned long long add=((unsigned long long)lo|((unsigned long long)hi<<32))&open;   // only open sectors matter
            if (add)
            {
              occ|=add;
              changed=true;
            }
          }
        }
        else
        {
          const uint32_t wq=__shfl_sync(WB_FULL,wedge,q);
          uint32_t cMaxA=0,cMinA=0xffffffffu,cMaxB=0,cMinB=0xffffffffu;
          if (in)
          {
            int s=WB_CL_FSECTOR?wb_sector64f(ddx,ddy):wb_sector64(ddx,ddy);
            int k1a=wq&255,k2a=(wq>>8)&255,k1b=(wq>>16)&255,k2b=wq>>24;
            if (s<0 || s==k1a || s==k2a || s==k1b || s==k2b)
            {
              uint32_t u;
              s=wb_sector64_exact(ddx,ddy,u);
              if (s==k1a) cMaxA=u;
              if (s==k2a) cMinA=u;
              if (s==k1b) cMaxB=u;
              if (s==k2b) cMinB=u;
            }
          }
          cMaxA=__reduce_max_sync(WB_FULL,cMaxA);
          cMinA=__reduce_min_sync(WB_FULL,cMinA);
          cMaxB=__reduce_max_sync(WB_FULL,cMaxB);
          cMinB=__reduce_min_sync(WB_FULL,cMinB);
          if (lane==q)
          {
            maxLowA=max(maxLowA,cMaxA);
            minHighA=min(minHighA,cMinA);
            maxLowB=max(maxLowB,cMaxB);
            minHighB=min(minHighB,cMinB);
          }
        }
      }
      if (pass==1)
      {
        // surrounded for sure once no empty run of 24 sectors is left
        if (!__any_sync(WB_FULL,changed))
          continue;                                 // nothing new for anybody: envelope and needs stand
        if (changed)
        {
          changed=false;
          open=wb_long_runs(occ);
          if (!open)
          {
            surrounded=true;
            done=true;
            live=false;
          }
#if WB_CL_XWANTS
          w.openq[lane]=open;
#endif
        }
#if WB_CL_XWANTS
        __syncwarp();
#endif
        liveMask=__ballot_sync(WB_FULL,live);
        if (!liveMask)
          break;
        // shrink the envelope when queries have finished; refresh the needed sectors
        if (liveMask!=envMask)
          envelope();
        needed();
#if WB_CL_REFILTER
        // The siblings still waiting in this entry: drop at once those no live query can want any more — the
        // necessary half of the test every pop makes (a reaching query that is still live, a sector some live
        // query still needs), for all 32 children in one step instead of one rejected pop each.
        if (!((w.wants[lane]&liveMask) && (w.cm[lane]&needAny)))
          w.keys[sp-1][lane]=0xffffffffu;
        __syncwarp();
#endif
      }
    }
    if (pass==1)
    {
      // undecided queries: empty run of >= 26 sectors -> not surrounded; 24..25 -> exact second walk
      if (!done)
      {
        unsigned long long empty=~occ;
        unsigned long long r24=wb_runs_ge(empty,24);
        if (r24 && !wb_runs_ge(empty,26))
        {
          // locate the runs: a run ends at sector e where r24 has a bit and sector e+1 is occupied
          int nrun=0;
          unsigned long long ends=r24&~wb_rotl64(empty,63);        // empty[e+1]==0  <=>  rotl(empty,63) bit e == 0
          while (ends && nrun<2)
          {
            int e=__ffsll((long long)ends)-1;
            ends&=ends-1;
            int len=0;
            while (len<64 && ((empty>>((e-len+64)&63))&1))
              len++;
            int k1=(e-len+64)&63,k2=(e+1)&63;
            if (nrun==0)
              wedge=(wedge&0xffff0000u)|(uint32_t)k1|((uint32_t)k2<<8);
            else
              wedge=(wedge&0x0000ffffu)|((uint32_t)k1<<16)|((uint32_t)k2<<24);
            wedgeMask|=(1ull<<k1)|(1ull<<k2);
            nrun++;
          }
        }
      }
    }
  }
finish:
  if (PASS==1)
  {
    // hand the undecided queries to the second pass
    if (have)
      wedgeBuf[me]=wedge;
    if (__any_sync(WB_FULL,wedgeMask!=0) && lane==0)
      chunkPending[chunk]=1;
  }
  if (PASS==2 && wedgeMask)
  {
    // exact gaps of the 24/25-sector runs; every other gap is < 25 sectors < 144 degrees
    bool gapA=(wedge&255)!=255 && ((minHighA-maxLowA)&0x7fffffffu)>=(uint32_t)WB_DEG144;
    bool gapB=((wedge>>16)&255)!=255 && ((minHighB-maxLowB)&0x7fffffffu)>=(uint32_t)WB_DEG144;
    surrounded=!(gapA || gapB);
  }
  if (PASS==1)
  {
    if (have && !foreign)
    {
      uint8_t lab;
      if (untiled)
        lab=clsIn[perm[me]];                         // never visited by classifyCylinder
      else
        lab=surrounded?1:2;                          // classify.cpp:158-162 (pending ones are overwritten by pass 2)
      labelSorted[me]=lab;
    }
    else if (have)
      labelSorted[me]=255;
  }
  else if (have && wedgeMask)
    labelSorted[me]=surrounded?1:2;
  unsigned mm=__ballot_sync(WB_FULL,margin);
  unsigned uu=__ballot_sync(WB_FULL,untiled);
  unsigned aa=__ballot_sync(WB_FULL,wedgeMask!=0);
  if (lane==0)
  {
    if (mm) atomicAdd(&counters[0],(unsigned long long)__popc(mm));
    if (uu) atomicAdd(&counters[1],(unsigned long long)__popc(uu));
    if (aa) atomicAdd(&counters[6],(unsigned long long)__popc(aa));
    atomicAdd(&counters[8],(unsigned long long)statNodes);
    atomicAdd(&counters[9],(unsigned long long)statChunks);
    atomicAdd(&counters[10],(unsigned long long)statPairs);
    atomicAdd(&counters[11],(unsigned long long)statNodes2);
    atomicAdd(&counters[12],(unsigned long long)statChunks2);
    atomicAdd(&counters[13],(unsigned long long)statPairs2);
    if (statNodes2) atomicAdd(&counters[14],1ull);
  }
}

// ---- classify order ------------------------------------------------------------------------------------------
// The label of a point does not depend on the order the kernel visits the store in, but its COST does: the warp's 32
// queries share one walk, and a chunk is opened for every query whose hyperboloid reaches its box.  In the canonical
// order (3-D Morton keys of the octree) 32 consecutive points of a surface hop between height cells and across the
// curve's jumps: boxes of 16.8 m2 on average on the bench tile, against 1.6 m2 for 32 points at 20 /m2.  Along a
// HILBERT curve over xy alone consecutive points, chunks and 32-chunk nodes are compact blobs (model on the 100 M
// tile: pair tests per warp 1014 -> 442, chunks opened 119 -> 68, nodes 251 -> 190).  So classify re-sorts the store
// by a 2 x 20-bit Hilbert index (5 radix passes), gathers coordinates, winner tile and input index in that order,
// builds the bounds hierarchy over it, walks it, and scatters the labels back to canonical order.
#ifndef WB_CL_HILBERT
#define WB_CL_HILBERT 1
#endif
__global__ void __launch_bounds__(256)
wb_hilbert_key_kernel(const double *__restrict__ sx,const double *__restrict__ sy,unsigned long long n,
                      double x0,double y0,double cellsPerUnit,unsigned long long *__restrict__ key,uint32_t *__restrict__ idx)
{
  unsigned long long j=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  if (j>=n)
    return;
  key[j]=wb_hilbert_index(sx[j],sy[j],x0,y0,cellsPerUnit);
  idx[j]=(uint32_t)j;
}

__global__ void __launch_bounds__(256)
wb_classify_gather_kernel(const uint32_t *__restrict__ ord,unsigned long long n,
                          const double *__restrict__ sx,const double *__restrict__ sy,const double *__restrict__ sz,
                          const uint32_t *__restrict__ winner,const uint32_t *__restrict__ perm,
                          double *__restrict__ hx,double *__restrict__ hy,double *__restrict__ hz,
                          uint32_t *__restrict__ hwinner,uint32_t *__restrict__ hperm)
{
  unsigned long long k=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  if (k>=n)
    return;
  const uint32_t j=ord[k];
  hx[k]=sx[j]; hy[k]=sy[j]; hz[k]=sz[j];
  hwinner[k]=winner[j];
  hperm[k]=perm[j];
}

__global__ void __launch_bounds__(256)
wb_classify_scatter_kernel(const uint32_t *__restrict__ ord,const uint8_t *__restrict__ hlabel,unsigned long long n,
                           uint8_t *__restrict__ labelSorted)
{
  unsigned long long k=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  if (k<n)
    labelSorted[ord[k]]=hlabel[k];
}

#if WB_CL_COMPACT2
__global__ void __launch_bounds__(256)
wb_pending_flag_kernel(const uint32_t *__restrict__ wedgeBuf,unsigned long long n,uint32_t *__restrict__ flag)
{
  unsigned long long i=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  if (i<n)
    flag[i]=wedgeBuf[i]!=0xffffffffu;
}

__global__ void __launch_bounds__(256)
wb_pending_scatter_kernel(const uint32_t *__restrict__ flag,const uint32_t *__restrict__ off,unsigned long long n,
                          uint32_t *__restrict__ list)
{
  unsigned long long i=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  if (i<n && flag[i])
    list[off[i]]=(uint32_t)i;
}
#endif

// ============================================================================ K3b: identical locations
// OctBuffer::put (octree.cpp:620-662) overwrites a stored point whose location equals the new one:
// of several points with one XYZ a single point stays in the store, at the position of the first
// inserted (lowest input index in canonical order).  Equal locations have equal keys, so the search
// stays inside a run of equal keys.  The survivors keep the canonical order; the others are moved
// behind them (key WB_KEY_DUP, stable re-sort) and later receive the survivor's label.
#define WB_KEY_DUP 0xfffffffffffffffeull

__global__ void __launch_bounds__(256)
wb_dup_find_kernel(const unsigned long long *__restrict__ keys,const double *__restrict__ sx,
                   const double *__restrict__ sy,const double *__restrict__ sz,unsigned long long nv,
                   uint32_t *__restrict__ flag,uint32_t *__restrict__ prev,unsigned long long *__restrict__ count)
{
  unsigned long long j=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  if (j>=nv)
    return;
  uint32_t f=0,p=0;
  if (j>0)
  {
    const unsigned long long k=keys[j];
    if (keys[j-1]==k)
    {
      const double x=sx[j],y=sy[j],z=sz[j];
      for (unsigned long long i=j;i>0 && keys[i-1]==k;)
      {
        i--;
        if (sx[i]==x && sy[i]==y && sz[i]==z)
        { // nearest earlier point at this location (may itself be a duplicate: resolved by wb_dup_jump)
          f=1;
          p=(uint32_t)i;
          break;
        }
      }
    }
  }
  flag[j]=f;
  prev[j]=p;
  if (f)
    atomicAdd(count,1ull);
}

__global__ void __launch_bounds__(256)
wb_dup_jump_kernel(const uint32_t *__restrict__ flag,uint32_t *prev,unsigned long long nv,
                   unsigned long long *__restrict__ changed)
// pointer jumping: prev[j] ends at the first point of its location (which is not flagged)
{
  unsigned long long j=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  if (j>=nv || !flag[j])
    return;
  uint32_t p=prev[j];
  if (flag[p])
  {
    prev[j]=prev[p];
    atomicAdd(changed,1ull);
  }
}

__global__ void __launch_bounds__(256)
wb_dup_mark_kernel(const uint32_t *__restrict__ flag,const uint32_t *__restrict__ prev,
                   const uint32_t *__restrict__ perm,unsigned long long nv,unsigned long long *keys,
                   uint32_t *__restrict__ dupIn,uint32_t *__restrict__ dupRep,unsigned long long *__restrict__ slot,
                   uint8_t *__restrict__ forcedIn,uint32_t ownFirst,uint32_t ownEnd)
{
  unsigned long long j=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  if (j>=nv || !flag[j])
    return;
  unsigned long long s=atomicAdd(slot,1ull);
  const uint32_t lost=perm[j],kept=perm[prev[j]];
  dupIn[s]=lost;
  dupRep[s]=kept;
  keys[j]=WB_KEY_DUP;
  // sharded run: one of OUR records lost to a halo point — that point's label is needed here
  if (forcedIn && lost>=ownFirst && lost<ownEnd && (kept<ownFirst || kept>=ownEnd))
    forcedIn[kept]=1;
}

__global__ void __launch_bounds__(256)
wb_dup_labels_kernel(const uint32_t *__restrict__ dupIn,const uint32_t *__restrict__ dupRep,unsigned long long n,
                     uint8_t *labelIn)
{
  unsigned long long u=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  if (u<n)
    labelIn[dupIn[u]]=labelIn[dupRep[u]];
}

// ============================================================================ K10: labels back to input order

__global__ void __launch_bounds__(256)
wb_scatter_labels_kernel(const uint8_t *__restrict__ labelSorted,const uint32_t *__restrict__ perm,
                         unsigned long long n,uint8_t *__restrict__ labelIn)
{
  unsigned long long j=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  if (j<n)
    labelIn[perm[j]]=labelSorted[j];
}

__global__ void __launch_bounds__(256)
wb_init_labels_kernel(const uint8_t *__restrict__ cls,unsigned long long n,uint8_t *__restrict__ labelIn)
{
  unsigned long long j=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  if (j<n)
    labelIn[j]=cls[j];
}

__global__ void __launch_bounds__(256)
wb_count_classes_kernel(const uint8_t *__restrict__ labelIn,const uint8_t *__restrict__ ret,unsigned long long n,
                        unsigned long long *__restrict__ counts)
// countClasses (threads.cpp:425-444): histogram of the class byte over the stored points
{
  __shared__ unsigned int h[256];
  h[threadIdx.x]=0;
  __syncthreads();
  for (unsigned long long j=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;j<n;
       j+=(unsigned long long)gridDim.x*blockDim.x)
    if (ret[j])
      atomicAdd(&h[labelIn[j]],1u);
  __syncthreads();
  if (h[threadIdx.x])
    atomicAdd(&counts[threadIdx.x],(unsigned long long)h[threadIdx.x]);
}

__global__ void __launch_bounds__(256)
wb_leaf_keys_kernel(const WbLeafDev *__restrict__ leaves,uint32_t nLeaves,const unsigned long long *__restrict__ keys,
                    unsigned long long *__restrict__ out)
{
  uint32_t i=blockIdx.x*blockDim.x+threadIdx.x;
  if (i<nLeaves)
    out[i]=keys[leaves[i].first];
}

__global__ void __launch_bounds__(256)
wb_copy_points_kernel(const int *__restrict__ x,const int *__restrict__ y,const int *__restrict__ z,
                      const uint8_t *__restrict__ c,unsigned long long n,
                      int *__restrict__ ox,int *__restrict__ oy,int *__restrict__ oz,uint8_t *__restrict__ oc,
                      uint8_t *__restrict__ oret)
{
  unsigned long long i=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  if (i>=n)
    return;
  ox[i]=x[i]; oy[i]=y[i]; oz[i]=z[i];
  oc[i]=c[i];
  if (oret)
    oret[i]=1;
}

__global__ void __launch_bounds__(256)
wb_test_math_kernel(const double *__restrict__ y,const double *__restrict__ x,unsigned long long n,
                    int *__restrict__ a,double *__restrict__ h,int *__restrict__ s)
{
  unsigned long long i=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  if (i>=n)
    return;
  a[i]=wb_atan2i(y[i],x[i]);
  h[i]=wb_hypot(x[i],y[i]);
  s[i]=wb_sector64(x[i],y[i]);
}
