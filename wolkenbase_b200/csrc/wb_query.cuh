// wb_query.cuh — K12: the OctStore query API on the device.
//
// OctStore::pointsIn / countPointsIn / hiLoPointsIn (octree.cpp:1214-1293) over the shapes of
// shape.cpp: Sphere, Paraboloid, Hyperboloid, Cylinder, Column.  The reference walks the octree
// (Octree::findBlocks 234-251 with Shape::intersect = in(closestPoint(cube))) and filters the
// candidate buckets with Shape::in; here a warp walks the 32-ary bounds hierarchy over the
// Morton-sorted points with a conservative version of the same closest-point test and applies the
// exact predicate (same operations, un-fused, glibc hypot) to the points of the chunks it reaches.
// Results are in canonical order, which is the reference's bucket order x in-bucket order.
#pragma once
#include <cstdint>

#define WB_SHAPE_SPHERE 0
#define WB_SHAPE_PARABOLOID 1
#define WB_SHAPE_HYPERBOLOID 2
#define WB_SHAPE_CYLINDER 3
#define WB_SHAPE_COLUMN 4

struct WbShapeDev { int type,pad; double p[6]; };
// sphere: cx,cy,cz,r | paraboloid: vx,vy,vz,radiusCurvature | hyperboloid: vx,vy,vz,r,slope
// cylinder: cx,cy,r | column: cx,cy,side

__device__ __forceinline__ bool wb_shape_in(const WbShapeDev &s,double x,double y,double z)
// Shape::in(xyz), shape.cpp:81-90 (Paraboloid), 127-135 (Hyperboloid), 175-178 (Sphere), 215-219 (Cylinder), 252-255 (Column)
{
  switch (s.type)
  {
    case WB_SHAPE_SPHERE:
      return wb_hypot(wb_hypot(__dsub_rn(x,s.p[0]),__dsub_rn(y,s.p[1])),__dsub_rn(z,s.p[2]))<=s.p[3];
    case WB_SHAPE_PARABOLOID:
    {
      const double d=wb_hypot(__dsub_rn(s.p[0],x),__dsub_rn(s.p[1],y)),zd=__dsub_rn(s.p[2],z),r=s.p[3];
      if (r==0)
        return d==0;
      const double q=__ddiv_rn(d,r);
      return __ddiv_rn(__dmul_rn(2.,zd),r)>=__dmul_rn(q,q);
    }
    case WB_SHAPE_HYPERBOLOID:
    {
      const double sl=s.p[4],por=__dmul_rn(s.p[3],__dmul_rn(sl,sl)),por2=__dmul_rn(por,por),cz=__dadd_rn(s.p[2],por);
      const double d=wb_hypot(__dsub_rn(s.p[0],x),__dsub_rn(s.p[1],y)),zd=__dsub_rn(cz,z),ds=__dmul_rn(d,sl);
      const double lhs=__dsub_rn(__dmul_rn(zd,zd),__dmul_rn(ds,ds));
      return (sl>0?zd>0:zd<0) && lhs>=por2;
    }
    case WB_SHAPE_CYLINDER:
      return wb_hypot(__dsub_rn(s.p[0],x),__dsub_rn(s.p[1],y))<=s.p[2];
    case WB_SHAPE_COLUMN:
      return fabs(__dsub_rn(s.p[0],x))<=s.p[2]/2 && fabs(__dsub_rn(s.p[1],y))<=s.p[2]/2;
  }
  return false;
}

__device__ __forceinline__ bool wb_shape_may_touch(const WbShapeDev &s,const WbBound &b)
// Conservative Shape::intersect for a box known by its xy extent and its lowest z (no highest z):
// the closest point of the box in xy, the most favourable z, and a relative slack of 1e-9.
{
  const double ax=s.p[0],ay=s.p[1];
  const double dx=ax<b.xmin?b.xmin-ax:(ax>b.xmax?ax-b.xmax:0.),dy=ay<b.ymin?b.ymin-ay:(ay>b.ymax?ay-b.ymax:0.);
  const double d2=dx*dx+dy*dy,slack=1e-9;
  switch (s.type)
  {
    case WB_SHAPE_SPHERE:
    {
      const double dz=s.p[2]<b.zmin?b.zmin-s.p[2]:0.,r=s.p[3];
      return d2+dz*dz<=r*r*(1+slack)+1e-300;
    }
    case WB_SHAPE_PARABOLOID:
    {
      const double r=s.p[3];
      if (r==0)
        return d2==0;
      if (r<0)
        return true;                                 // opens upward: the box's top is unknown
      const double zd=s.p[2]-b.zmin;                   // largest zdist any point of the box can have
      return 2*zd*r*(1+(zd>0?slack:-slack))+1e-300>=d2*(1-slack);
    }
    case WB_SHAPE_HYPERBOLOID:
    {
      const double sl=s.p[4];
      if (!(sl>0))
        return true;
      const double por=s.p[3]*sl*sl,zd=s.p[2]+por-b.zmin;
      return zd>0 && zd*zd*(1+slack)>=(d2*sl*sl+por*por)*(1-slack);
    }
    case WB_SHAPE_CYLINDER:
      return d2<=s.p[2]*s.p[2]*(1+slack)+1e-300;
    case WB_SHAPE_COLUMN:
      return dx<=s.p[2]/2*(1+slack) && dy<=s.p[2]/2*(1+slack);
  }
  return false;
}

#define WB_Q_WARPS 4
#define WB_Q_STACK 256

__global__ void __launch_bounds__(WB_Q_WARPS*32)
wb_query_kernel(const WbShapeDev *__restrict__ shapes,unsigned long long nq,
                const double *__restrict__ sx,const double *__restrict__ sy,const double *__restrict__ sz,
                unsigned long long nv,const WbBound *__restrict__ bounds,const uint32_t *__restrict__ levelOff,
                const uint32_t *__restrict__ levelCnt,int nLevels,
                unsigned long long *__restrict__ count,double *__restrict__ lo,double *__restrict__ hi)
// one warp per shape: countPointsIn and hiLoPointsIn together
{
  __shared__ uint32_t stackS[WB_Q_WARPS][WB_Q_STACK];    // entry = level<<28 | index
  const int warp=threadIdx.x>>5,lane=threadIdx.x&31;
  const unsigned long long q=(unsigned long long)blockIdx.x*WB_Q_WARPS+warp;
  if (q>=nq)
    return;
  const WbShapeDev s=shapes[q];
  uint32_t *stack=stackS[warp];
  int sp=0;
  {
    const int top=nLevels-1;
    const uint32_t cnt=levelCnt[top];                   // <= 32
    bool t=lane<cnt && wb_shape_may_touch(s,bounds[levelOff[top]+lane]);
    uint32_t m=__ballot_sync(0xffffffffu,t);
    if (t)
      stack[__popc(m&((1u<<lane)-1))]=((uint32_t)top<<28)|(uint32_t)lane;
    sp=__popc(m);
  }
  __syncwarp();
  unsigned long long n=0;
  double mn=INFINITY,mx=-INFINITY;
  while (sp>0)
  {
    const uint32_t e=stack[--sp];
    __syncwarp();
    const int level=(int)(e>>28);
    const uint32_t node=e&0x0fffffffu;
    if (level==0)
    {
      const unsigned long long j=(unsigned long long)node*32+lane;
      if (j<nv)
      {
        const double x=sx[j],y=sy[j],z=sz[j];
        if (wb_shape_in(s,x,y,z))
        {
          n++;
          mn=fmin(mn,z);
          mx=fmax(mx,z);
        }
      }
    }
    else
    {
      const uint32_t c=node*32+lane;
      bool t=c<levelCnt[level-1] && wb_shape_may_touch(s,bounds[levelOff[level-1]+c]);
      uint32_t m=__ballot_sync(0xffffffffu,t);
      // children are pushed in reverse so that they pop in ascending (canonical) order
      if (t)
        stack[sp+__popc(m)-1-__popc(m&((1u<<lane)-1))]=((uint32_t)(level-1)<<28)|c;
      sp+=__popc(m);
      __syncwarp();
    }
  }
  #pragma unroll
  for (int o=16;o;o>>=1)
  {
    n+=__shfl_xor_sync(0xffffffffu,n,o);
    mn=fmin(mn,__shfl_xor_sync(0xffffffffu,mn,o));
    mx=fmax(mx,__shfl_xor_sync(0xffffffffu,mx,o));
  }
  if (lane==0)
  {
    if (count) count[q]=n;
    if (lo) lo[q]=mn;
    if (hi) hi[q]=mx;
  }
}

__global__ void __launch_bounds__(256)
wb_query_flag_kernel(WbShapeDev s,const double *__restrict__ sx,const double *__restrict__ sy,
                     const double *__restrict__ sz,unsigned long long nv,uint32_t *__restrict__ flag)
// pointsIn for one shape: the exact predicate on every stored point (12 B/point of HBM traffic)
{
  unsigned long long j=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  if (j<nv)
    flag[j]=wb_shape_in(s,sx[j],sy[j],sz[j])?1u:0u;
}

__global__ void __launch_bounds__(256)
wb_query_emit_kernel(const uint32_t *__restrict__ flag,const uint32_t *__restrict__ pos,unsigned long long nv,
                     const uint32_t *__restrict__ perm,const double *__restrict__ sx,const double *__restrict__ sy,
                     const double *__restrict__ sz,unsigned long long cap,uint32_t *__restrict__ outPos,
                     uint32_t *__restrict__ outIdx,double *__restrict__ ox,double *__restrict__ oy,double *__restrict__ oz)
{
  unsigned long long j=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  if (j<nv && flag[j])
  {
    const uint32_t o=pos[j];
    if (o<cap)
    {
      outPos[o]=(uint32_t)j;
      outIdx[o]=perm[j];
      ox[o]=sx[j]; oy[o]=sy[j]; oz[o]=sz[j];
    }
  }
}
