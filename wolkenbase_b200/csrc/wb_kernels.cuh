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

#define WB_FULL 0xffffffffu

struct WbBound { double xmin,xmax,ymin,ymax,zmin; };

struct WbSegment            // one wb_add_las call
{
  unsigned long long first,count;
  double scale[3],offset[3],unit;
};
#define WB_MAX_SEGMENTS 256
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
  __shared__ uint32_t sw[(WB_DEC_THREADS*WB_DEC_MAXLEN+32)/4+2];
  const unsigned long long first=(unsigned long long)blockIdx.x*WB_DEC_THREADS;
  const unsigned long long cnt=n-first<WB_DEC_THREADS?n-first:WB_DEC_THREADS;
  const unsigned long long b0=first*recLen,b1=(first+cnt)*recLen;
  const uintptr_t p0=(uintptr_t)recs+b0;
  const uintptr_t a0=p0&~(uintptr_t)15;               // aligned start (may precede our span)
  const uint32_t head=(uint32_t)(p0-a0);
  const uint32_t total=head+(uint32_t)(b1-b0);
  const uint32_t nvec=(total+15)/16;
  const uintptr_t endAll=(uintptr_t)recs+n*(unsigned long long)recLen;
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

// ============================================================================ K2: Morton keys

__device__ __forceinline__ unsigned long long wb_morton(double x,double y,double z,
                                                        double cx,double cy,double cz,double side)
// 21 steps of Octree::findBlock's descent: bit = coordinate >= centre, child = z*4+y*2+x
// (octree.cpp:199-216); child centre = centre +- side/4 (octree.cpp:335-337).  All centre
// arithmetic is exact (dyadic), so no rounding mode matters here.
{
  unsigned long long key=0;
  double q=side*0.25;
  #pragma unroll
  for (int l=0;l<21;l++)
  {
    int xb=x>=cx,yb=y>=cy,zb=z>=cz;
    key=(key<<3)|(unsigned long long)(zb*4+yb*2+xb);
    cx+=xb?q:-q;
    cy+=yb?q:-q;
    cz+=zb?q:-q;
    q*=0.5;
  }
  return key;
}

__global__ void __launch_bounds__(256)
wb_keygen_kernel(const int *__restrict__ xi,const int *__restrict__ yi,const int *__restrict__ zi,
                 const uint8_t *__restrict__ ret,unsigned long long first,unsigned long long cnt,
                 WbSegment seg,double cx,double cy,double cz,double side,
                 unsigned long long *__restrict__ key,uint32_t *__restrict__ idx)
{
  unsigned long long t=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  if (t>=cnt)
    return;
  unsigned long long i=first+t;
  double x=wb_coord(seg.offset[0],seg.scale[0],xi[i],seg.unit);
  double y=wb_coord(seg.offset[1],seg.scale[1],yi[i],seg.unit);
  double z=wb_coord(seg.offset[2],seg.scale[2],zi[i],seg.unit);
  unsigned long long k=wb_morton(x,y,z,cx,cy,cz,side);
  if (ret[i]==0)
    k=~0ull;                                          // dropped record: sorts behind every real key
  key[i]=k;
  idx[i]=(uint32_t)i;
}

// ============================================================================ K3: gather to canonical order

__global__ void __launch_bounds__(256)
wb_gather_kernel(const uint32_t *__restrict__ perm,unsigned long long n,
                 const int *__restrict__ xi,const int *__restrict__ yi,const int *__restrict__ zi,
                 const WbSegments *__restrict__ segs,
                 double *__restrict__ sx,double *__restrict__ sy,double *__restrict__ sz)
{
  unsigned long long j=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  if (j>=n)
    return;
  uint32_t i=perm[j];
  int lo=0,hi=segs->n-1;
  while (lo<hi)
  {
    int mid=(lo+hi+1)>>1;
    if (segs->s[mid].first<=i)
      lo=mid;
    else
      hi=mid-1;
  }
  const WbSegment &s=segs->s[lo];
  sx[j]=wb_coord(s.offset[0],s.scale[0],xi[i],s.unit);
  sy[j]=wb_coord(s.offset[1],s.scale[1],yi[i],s.unit);
  sz[j]=wb_coord(s.offset[2],s.scale[2],zi[i],s.unit);
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
                      unsigned long long *__restrict__ pairKey,uint32_t *__restrict__ pairVal)
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
wb_segment_kernel(const unsigned long long *__restrict__ pairKey,unsigned long long m,
                  uint32_t *__restrict__ tStart,uint32_t *__restrict__ tCount)
{
  unsigned long long i=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  if (i>=m)
    return;
  unsigned long long t=pairKey[i];
  if (i==0 || pairKey[i-1]!=t)
    tStart[t]=(uint32_t)i;
  if (i+1==m || pairKey[i+1]!=t)
    tCount[t]=(uint32_t)i+1;      // end for now; turned into a count by the scan kernel
}

// ============================================================================ K7: tile scan
// scanCylinder (scan.cpp:31-140), one thread per non-empty tile, streaming over the tile's
// points in canonical order.  Pairwise sums keep the reference's association
// (manysum.cpp:120-154): aligned power-of-two blocks summed as perfect trees, merged like a
// binary counter, block totals added from the smallest block up.

struct WbPairwise
{
  double lv[28];
  __device__ void push(double v,uint32_t i)       // i = index of v (0-based)
  {
    // carry: merge with every level whose bit is set in i (they are complete blocks)
    int l=0;
    uint32_t m=i;
    while (m&1)
    {
      v=__dadd_rn(lv[l],v);
      m>>=1;
      l++;
    }
    lv[l]=v;
  }
  __device__ double total(uint32_t n) const
  {
    double s=0;
    for (int l=0;l<28;l++)
      if ((n>>l)&1)
        s=__dadd_rn(s,lv[l]);
    return s;
  }
};

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

__global__ void __launch_bounds__(64)
wb_scan_kernel(const uint32_t *__restrict__ tStart,uint32_t *__restrict__ tCount,uint32_t nTiles,
               const uint32_t *__restrict__ pairVal,
               const double *__restrict__ sx,const double *__restrict__ sy,const double *__restrict__ sz,
               WbSnake snake,double minHyp,
               int *__restrict__ tNPoints,uint8_t *__restrict__ tTree,double *__restrict__ tDensity,
               double *__restrict__ tHyp,double *__restrict__ tHeight,unsigned long long *nNonEmpty)
{
  uint32_t t=blockIdx.x*blockDim.x+threadIdx.x;
  if (t>=nTiles)
    return;
  uint32_t end=tCount[t];
  if (end==0)
  {
    tNPoints[t]=0;
    return;
  }
  uint32_t start=tStart[t],cnt=end-start;
  tCount[t]=cnt;
  atomicAdd(nNonEmpty,1ull);
  int ex,ey;
  wb_to_flowsnake((int)t+snake.lo,ex,ey);
  double ccx,ccy;
  wb_tile_center(ex,ey,snake,ccx,ccy);
  // normal equations: sums of x*x, x*y, y*y, x*1, y*1, x*z, y*z, 1*z  (1*1 sums to cnt exactly)
  WbPairwise pxx,pxy,pyy,px,py,pxz,pyz,pz;
  for (uint32_t i=0;i<cnt;i++)
  {
    uint32_t k=pairVal[start+i];
    double x=__dsub_rn(sx[k],ccx),y=__dsub_rn(sy[k],ccy),z=sz[k];
    pxx.push(__dmul_rn(x,x),i);
    pxy.push(__dmul_rn(y,x),i);      // mt[1][k]*mt[0][k]
    pyy.push(__dmul_rn(y,y),i);
    px.push(x,i);                    // 1*x is exact
    py.push(y,i);
    pxz.push(__dmul_rn(x,z),i);
    pyz.push(__dmul_rn(y,z),i);
    pz.push(z,i);
  }
  WbMat3 m;
  m.a[0][0]=pxx.total(cnt);
  m.a[1][0]=m.a[0][1]=pxy.total(cnt);
  m.a[1][1]=pyy.total(cnt);
  m.a[2][0]=m.a[0][2]=px.total(cnt);
  m.a[2][1]=m.a[1][2]=py.total(cnt);
  m.a[2][2]=(double)cnt;
  m.b[0]=pxz.total(cnt);
  m.b[1]=pyz.total(cnt);
  m.b[2]=pz.total(cnt);
  wb_gausselim(m);
  double sl0=m.a[0][0]==0?NAN:m.b[0],sl1=m.a[1][1]==0?NAN:m.b[1];
  double len=wb_hypot(sl0,sl1);
  if (len>1)
  {
    sl0=__ddiv_rn(sl0,len);
    sl1=__ddiv_rn(sl1,len);
  }
  if (isnan(sl0) || isnan(sl1))
    sl0=sl1=0;
  double bottom=INFINITY,bottom2=INFINITY,top=-INFINITY;
  for (uint32_t i=0;i<cnt;i++)
  {
    uint32_t k=pairVal[start+i];
    double x=__dsub_rn(sx[k],ccx),y=__dsub_rn(sy[k],ccy);
    double zt=__dadd_rn(__dmul_rn(sl1,y),__dmul_rn(sl0,x));      // dot(): a.y*b.y+a.x*b.x
    double zu=__dsub_rn(sz[k],zt);
    if (zu<bottom)
    {
      bottom2=bottom;
      bottom=zu;
    }
    if (zu>top)
      top=zu;
  }
  if (isinf(bottom2))
    bottom2=bottom;
  int histo[7]={0,0,0,0,0,0,0};
  uint32_t nBottom=0;
  const double cut=__dadd_rn(bottom2,__dmul_rn(2.0,snake.radius));
  const double rin=__ddiv_rn(snake.radius,WB_SQRT7);
  for (uint32_t i=0;i<cnt;i++)
  {
    uint32_t k=pairVal[start+i];
    double x=__dsub_rn(sx[k],ccx),y=__dsub_rn(sy[k],ccy);
    double zt=__dadd_rn(__dmul_rn(sl1,y),__dmul_rn(sl0,x));
    double zu=__dsub_rn(sz[k],zt);
    if (zu<cut)
    {
      int sector=(int)wb_lrint(__ddiv_rn(__dmul_rn(atan2(y,x),3.0),WB_PI));
      if (sector<0)
        sector+=6;
      sector=(sector%6)+1;
      if (wb_hypot(x,y)<rin)
        sector=0;
      histo[sector]++;
      nBottom++;
    }
  }
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

// ============================================================================ K8: postscan

__global__ void __launch_bounds__(128)
wb_postscan_kernel(const int *__restrict__ tNPoints,const uint8_t *__restrict__ tTree,uint32_t nTiles,
                   WbSnake snake,double *__restrict__ tHyp)
// postscanCylinder (scan.cpp:142-179).  Reads only nPoints/treeFlags of other tiles, writes only
// this tile's hyperboloidSize: no ordering hazard.  A tile outside the table has nPoints == 0.
{
  uint32_t t=blockIdx.x*blockDim.x+threadIdx.x;
  if (t>=nTiles || tNPoints[t]==0)
    return;
  const int rx[6]={1,1,0,-1,-1,0},ry[6]={0,1,1,0,-1,-1};       // root1, eisenstein.cpp:51
  int ex,ey,count=0;
  wb_to_flowsnake((int)t+snake.lo,ex,ey);
  if (tTree[t]&1)
  {
    int i=1,ringcount,nontree;
    do
    {
      ringcount=nontree=0;
      for (int j=0;j<6;j++)
      {
        long long n;
        if (!wb_from_flowsnake(ex+rx[j]*i,ey+ry[j]*i,n) || n<snake.lo || n>snake.hi)
          continue;
        uint32_t o=(uint32_t)(n-snake.lo);
        if (tNPoints[o])
        {
          ringcount++;
          if (tTree[o]&1)
            count++;
          else
            nontree++;
        }
      }
      ++i;
    } while (ringcount && !nontree);
  }
  double h=tHyp[t],c=__ddiv_rn(__dmul_rn((double)count,snake.spacing),6.0);
  tHyp[t]=sqrt(__dadd_rn(__dmul_rn(h,h),__dmul_rn(c,c)));
}

// ============================================================================ K9: classify
// One warp owns one chunk of 32 consecutive points of the canonical order (spatial neighbours)
// as its 32 QUERY points, one per lane.  The warp walks the bucket hierarchy once for all 32:
// a node is entered if the downward hyperboloid of ANY of its queries can reach the node's
// lowest point at the node's nearest xy; the same test per lane decides which queries look at a
// chunk.  Candidate chunks are staged in shared memory and broadcast to the lanes.  Points found
// inside a query's hyperboloid go to a per-warp queue of (query, dx, dy); whenever 32 are
// queued, all lanes compute one atan2i each and fold the bearings into per-query angular bins
// (16 bins of 22.5 degrees, min and max bearing per bin) from which "no gap >= 144 degrees" is
// decided exactly (surround(), classify.cpp:67-94): gaps inside a bin are < 22.5 degrees, and
// gaps between bins are max-of-previous-bin to min-of-next-bin.

#define WB_CL_WARPS 8
#define WB_QCAP 64

struct WbClassifyWarp
{
  double qx[32],qy[32],qcz[32],qpor2[32];
  double cx[32],cy[32],cz[32];
  double pdx[WB_QCAP],pdy[WB_QCAP];
  int pq[WB_QCAP];
  uint32_t bmin[16][32],bmax[16][32];
  uint32_t qmask[32],touched[32];
  uint32_t stBase[8],stMask[8];
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

__device__ __forceinline__ bool wb_surrounded(const WbClassifyWarp &w,int lane)
{
  uint32_t mask=w.qmask[lane];
  if (mask==0)
    return false;
  int f=__ffs(mask)-1;
  uint32_t firstMin=w.bmin[f][lane],prevMax=w.bmax[f][lane];
  bool distinct=firstMin!=prevMax || (mask&(mask-1));
  if (!distinct)
    return false;
  bool ok=true;
  uint32_t m=mask&(mask-1);
  while (m)
  {
    int b=__ffs(m)-1;
    m&=m-1;
    if (w.bmin[b][lane]-prevMax>=(uint32_t)WB_DEG144)
      ok=false;
    prevMax=w.bmax[b][lane];
  }
  if (((firstMin-prevMax)&0x7fffffffu)>=(uint32_t)WB_DEG144)
    ok=false;
  return ok;
}

__device__ __forceinline__ void wb_drain(WbClassifyWarp &w,int lane,int cnt)
// lanes 0..cnt-1 each turn one queued (query,dx,dy) into a bearing and fold it into the bins
{
  if (lane<cnt)
  {
    int q=w.pq[lane];
    int a=wb_atan2i(w.pdy[lane],w.pdx[lane]);        // dir(P,Q)=atan2i(Q-P), point.cpp:194-197
    uint32_t u=(uint32_t)a&0x7fffffffu;
    uint32_t b=u>>27;
    atomicMin(&w.bmin[b][q],u);
    atomicMax(&w.bmax[b][q],u);
    atomicOr(&w.qmask[q],1u<<b);
    w.touched[q]=1;
  }
  __syncwarp();
}

__global__ void __launch_bounds__(WB_CL_WARPS*32)
wb_classify_kernel(const double *__restrict__ sx,const double *__restrict__ sy,const double *__restrict__ sz,
                   unsigned long long n,uint32_t nChunks,
                   const WbBound *__restrict__ bounds,const uint32_t *__restrict__ levelOff,
                   const uint32_t *__restrict__ levelCnt,int nLevels,
                   const uint32_t *__restrict__ winner,const double *__restrict__ tHyp,
                   double maxSlope,double thickness,
                   const uint8_t *__restrict__ clsIn,const uint32_t *__restrict__ perm,
                   uint8_t *__restrict__ labelSorted,unsigned long long *__restrict__ counters)
{
  extern __shared__ __align__(16) unsigned char smraw[];
  WbClassifyWarp &w=reinterpret_cast<WbClassifyWarp *>(smraw)[threadIdx.x>>5];
  const int lane=threadIdx.x&31;
  const uint32_t chunk=blockIdx.x*WB_CL_WARPS+(threadIdx.x>>5);
  if (chunk>=nChunks)
    return;
  const unsigned long long me=(unsigned long long)chunk*32+lane;
  const bool have=me<n;
  const double s2=maxSlope*maxSlope;
  double px=0,py=0,pcz=-INFINITY,ppor2=INFINITY;
  bool done=true,untiled=false;
  if (have)
  {
    px=sx[me];
    py=sy[me];
    uint32_t wt=winner[me];
    if (wt!=0xffffffffu)
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
  w.qmask[lane]=0; w.touched[lane]=0;
  #pragma unroll
  for (int b=0;b<16;b++)
  {
    w.bmin[b][lane]=0xffffffffu;
    w.bmax[b][lane]=0;
  }
  // group envelope: xy box of the queries, highest centre, smallest polar radius
  double gx0=have&&!done?px:INFINITY,gx1=have&&!done?px:-INFINITY;
  double gy0=have&&!done?py:INFINITY,gy1=have&&!done?py:-INFINITY;
  double gcz=pcz,gpor2=ppor2;
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
  __syncwarp();
  int sp=0,qcount=0;
  bool margin=false;
  if (!__all_sync(WB_FULL,done))
  {
    // root level: up to 32 nodes
    const int top=nLevels-1;
    uint32_t cntTop=levelCnt[top];
    bool pass=false;
    if ((uint32_t)lane<cntTop)
      pass=wb_reach(gx0,gx1,gy0,gy1,gcz,gpor2,s2,bounds[levelOff[top]+lane]);
    uint32_t m=__ballot_sync(WB_FULL,pass);
    if (m)
    {
      if (lane==0)
      {
        w.stLevel[0]=top;
        w.stBase[0]=0;
        w.stMask[0]=m;
      }
      sp=1;
    }
    __syncwarp();
  }
  while (sp>0)
  {
    // pop the lowest child of the top entry
    int level=w.stLevel[sp-1];
    uint32_t base=w.stBase[sp-1],mask=w.stMask[sp-1];
    int bit=__ffs(mask)-1;
    uint32_t node=base+bit;
    __syncwarp();
    mask&=mask-1;
    if (mask)
    {
      if (lane==0)
        w.stMask[sp-1]=mask;
    }
    else
      sp--;
    __syncwarp();
    if (level>0)
    {
      // test the node's 32 children against the group envelope
      uint32_t c=node*32+lane,cc=levelCnt[level-1];
      bool pass=false;
      if (c<cc)
        pass=wb_reach(gx0,gx1,gy0,gy1,gcz,gpor2,s2,bounds[levelOff[level-1]+c]);
      uint32_t m=__ballot_sync(WB_FULL,pass);
      if (m)
      {
        if (lane==0)
        {
          w.stLevel[sp]=level-1;
          w.stBase[sp]=node*32;
          w.stMask[sp]=m;
        }
        sp++;
      }
      __syncwarp();
      continue;
    }
    // ---- a candidate chunk: which of my queries can reach it?
    WbBound cb=bounds[node];                        // level 0 offset is 0
    bool lanePass=!done && wb_reach(px,px,py,py,pcz,ppor2,s2,cb);
    if (!__any_sync(WB_FULL,lanePass))
      continue;
    {
      unsigned long long j=(unsigned long long)node*32+lane;
      bool ok=j<n;
      w.cx[lane]=ok?sx[j]:0.0;
      w.cy[lane]=ok?sy[j]:0.0;
      w.cz[lane]=ok?sz[j]:INFINITY;
    }
    __syncwarp();
    #pragma unroll 2
    for (int p=0;p<32;p++)
    {
      bool in=false;
      double dx=0,dy=0;
      if (lanePass)
      {
        double zd=__dsub_rn(pcz,w.cz[p]);          // centre.z - pnt.z, shape.cpp:130
        if (zd>0)
        {
          dx=__dsub_rn(px,w.cx[p]);
          dy=__dsub_rn(py,w.cy[p]);
          double zz=zd*zd,hs2=(dx*dx+dy*dy)*s2;
          double diff=zz-hs2-ppor2,tol=1e-13*(zz+hs2+ppor2);
          if (diff>tol)
            in=true;
          else if (diff>=-tol)
          {
            // too close to call without the reference's exact expression (shape.cpp:127-135)
            double d=wb_hypot(dx,dy);
            double ds=__dmul_rn(d,maxSlope);
            double lhs=__dsub_rn(__dmul_rn(zd,zd),__dmul_rn(ds,ds));
            in=lhs>=ppor2;
            if (fabs(lhs-ppor2)<=1e-12*(zz+ppor2) && (dx!=0 || dy!=0))
              margin=true;
          }
          in=in && (dx!=0 || dy!=0);               // dist(...) != 0, classify.cpp:150
        }
      }
      uint32_t bm=__ballot_sync(WB_FULL,in);
      if (bm)
      {
        if (in)
        {
          int slot=qcount+__popc(bm&((1u<<lane)-1));
          w.pq[slot]=lane;
          w.pdx[slot]=-dx;                         // Q-P
          w.pdy[slot]=-dy;
        }
        qcount+=__popc(bm);
        __syncwarp();
        if (qcount>=32)
        {
          wb_drain(w,lane,32);
          // move the tail down
          int rest=qcount-32;
          int tq=0; double tx=0,ty=0;
          if (lane<rest)
          {
            tq=w.pq[32+lane]; tx=w.pdx[32+lane]; ty=w.pdy[32+lane];
          }
          __syncwarp();
          if (lane<rest)
          {
            w.pq[lane]=tq; w.pdx[lane]=tx; w.pdy[lane]=ty;
          }
          qcount=rest;
          __syncwarp();
          if (!done && w.touched[lane])
          {
            w.touched[lane]=0;
            if (wb_surrounded(w,lane))
            {
              done=true;
              lanePass=false;
            }
          }
        }
      }
    }
    __syncwarp();
    if (__all_sync(WB_FULL,done))
      break;
  }
  // Entries still queued for finished queries are harmless: extra bearings can only split gaps.
  if (qcount>0)
    wb_drain(w,lane,qcount);
  __syncwarp();
  if (have)
  {
    uint8_t lab;
    if (untiled)
      lab=clsIn[perm[me]];                           // never visited by classifyCylinder
    else
      lab=(done || wb_surrounded(w,lane))?1:2;       // classify.cpp:158-162
    labelSorted[me]=lab;
  }
  unsigned mm=__ballot_sync(WB_FULL,margin && have);
  unsigned uu=__ballot_sync(WB_FULL,untiled);
  if (lane==0)
  {
    if (mm) atomicAdd(&counters[0],(unsigned long long)__popc(mm));
    if (uu) atomicAdd(&counters[1],(unsigned long long)__popc(uu));
  }
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
wb_compact_tiles_kernel(const int *__restrict__ tNPoints,uint32_t nTiles,uint32_t *__restrict__ flag)
{
  uint32_t t=blockIdx.x*blockDim.x+threadIdx.x;
  if (t<nTiles)
    flag[t]=tNPoints[t]!=0;
}
