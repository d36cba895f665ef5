// wb_device.cuh — device-side arithmetic that must match the reference bit for bit.
//
// nvcc contracts a*b+c into FMA by default; the reference is compiled without FMA, so every
// expression whose rounding matters is spelled with __dmul_rn/__dadd_rn/__dsub_rn (never
// contracted).  Division and sqrt are IEEE-correct on the device.
#pragma once
#include <cstdint>
#include <climits>
#include <cuda_runtime.h>

#define WB_DEG30  0x0aaaaaab   // angle.h:105
#define WB_DEG45  0x10000000
#define WB_DEG90  0x20000000
#define WB_DEG144 0x33333333   // angle.h:115
#define WB_DEG180 0x40000000
#define WB_SQRT_3_4 0.86602540378443864676372317   // eisenstein.h:33
#define WB_SQRT7 2.6457513110645905905016          // scan.h:23
#define WB_PI 3.14159265358979323846

struct WbSnake            // Flowsnake state (flowsnake.h:96-104) + derived radius
{
  double spacing,ccx,ccy,radius;
  int lo,hi;
};

struct WbParams
{
  double tileSize,maxSlope,thickness,minHyp;
};

// Lookup tables live in global memory and are read through the read-only path: lanes index them
// with different addresses, which constant memory would serialise.
__device__ double g_tanTable[512];
__device__ double g_cosTable[512];
__device__ double g_sinTable[512];
__device__ unsigned char g_fwdTable[96];     // [6][8], flowsnake.cpp:46-54 padded to 8 columns, then its inverse (wb_host.h)

// ---- exact helpers ---------------------------------------------------------------------

__device__ __forceinline__ double wb_coord(double off,double scale,int v,double unit)
// las.cpp:808: (offset + scale*int)*unit, un-fused
{
  return __dmul_rn(__dadd_rn(off,__dmul_rn(scale,(double)v)),unit);
}

__device__ __forceinline__ void wb_two_prod(double a,double b,double &p,double &e)
{
  p=__dmul_rn(a,b);
  e=__fma_rn(a,b,-p);
}

__device__ __forceinline__ void wb_two_sum(double a,double b,double &s,double &e)
{
  s=__dadd_rn(a,b);
  double bb=__dsub_rn(s,a);
  e=__dadd_rn(__dsub_rn(a,__dsub_rn(s,bb)),__dsub_rn(b,bb));
}

__device__ __noinline__ double wb_hypot(double x,double y)
// hypot() as libm gives it to dist() (point.cpp:189-192).  glibc >= 2.35 (the image has 2.39)
// computes h = sqrt(x^2+y^2) and applies one error-compensating step (Borges' algorithm, the
// non-FMA branch of sysdeps/ieee754/dbl-64/e_hypot.c); restating that sequence with un-fused
// operations reproduces libm bit for bit (checked on 2e5 random pairs against the host libm;
// it is NOT always the correctly rounded value).  The scaling branches for huge/tiny inputs
// are unreachable for coordinates in metres.
{
  x=fabs(x);
  y=fabs(y);
  double ax=x<y?y:x,ay=x<y?x:y;
  if (__dmul_rn(ax,2.220446049250313e-16)>=ay)
    return __dadd_rn(ax,ay);
  double h=sqrt(__dadd_rn(__dmul_rn(ax,ax),__dmul_rn(ay,ay)));
  double t1,t2;
  if (h<=__dmul_rn(2.0,ay))
  {
    double delta=__dsub_rn(h,ay);
    t1=__dmul_rn(ax,__dsub_rn(__dmul_rn(2.0,delta),ax));
    t2=__dmul_rn(__dsub_rn(delta,__dmul_rn(2.0,__dsub_rn(ax,ay))),delta);
  }
  else
  {
    double delta=__dsub_rn(h,ax);
    t1=__dmul_rn(__dmul_rn(2.0,delta),__dsub_rn(ax,__dmul_rn(2.0,ay)));
    t2=__dadd_rn(__dmul_rn(__dsub_rn(__dmul_rn(4.0,delta),ay),ay),__dmul_rn(delta,delta));
  }
  return __dsub_rn(h,__ddiv_rn(__dadd_rn(t1,t2),__dmul_rn(2.0,h)));
}

__device__ __forceinline__ long long wb_lrint(double v)
{
  return __double2ll_rn(v);
}

__device__ __noinline__ int wb_atan2i(double y,double x)
// atan2i(): angle.cpp:117-155 (octant folding, 9-step bisection on the tangent table, rotation by
// the bin centre, cubic correction).
{
  int ret=0,h;
  double t,nx;
  if (x<0)
  {
    ret+=(y>0)?WB_DEG180:-WB_DEG180;
    y=-y;
    x=-x;
  }
  if (y>x)
  {
    ret+=WB_DEG90;
    t=x; x=y; y=-t;
  }
  if (-y>x)
  {
    ret-=WB_DEG90;
    t=x; x=-y; y=t;
  }
  t=__ddiv_rn(y,x);
  #pragma unroll
  for (h=WB_DEG45/2;h>WB_DEG45/1024;h/=2)
    if (t>__ldg(&g_tanTable[(((ret+WB_DEG45)&0x1ff00000)>>20)-1]))
      ret+=h;
    else
      ret-=h;
  h=511-(((ret+WB_DEG45)&0x1ff00000)>>20);
  double ch=__ldg(&g_cosTable[h]),sh=__ldg(&g_sinTable[h]);
  nx=__dsub_rn(__dmul_rn(x,ch),__dmul_rn(y,sh));
  y=__dadd_rn(__dmul_rn(y,ch),__dmul_rn(x,sh));
  x=nx;
  {
    double q=__ddiv_rn(y,x);
    double c=__dmul_rn(__dmul_rn(q,q),q);
    // 0x40000000/M_PIl*y/x - 1.1392738508503886e8*c: the reference evaluates the first term in
    // long double.  Here it is carried in double-double (K = the long double constant split in
    // two), which agrees with the 64-bit-mantissa result unless the value is within ~1e-13 of a
    // half-integer.
    const double kh=341782637.7882158,kl=-2.112938091158867e-08;
    double ph,pl;
    wb_two_prod(kh,y,ph,pl);
    pl=__fma_rn(kl,y,pl);
    double qh=__ddiv_rn(ph,x);
    double ql=__ddiv_rn(__dadd_rn(__fma_rn(-qh,x,ph),pl),x);
    double term=__dmul_rn(1.1392738508503886e8,c);
    double sh,se;
    wb_two_sum(qh,-term,sh,se);
    double r=rint(sh);
    double d=__dadd_rn(__dsub_rn(sh,r),__dadd_rn(se,ql));
    if (d>0.5)
      r+=1.0;
    else if (d<-0.5)
      r-=1.0;
    ret+=(int)r;
  }
  if (x==0 && y==0)
    ret=0;
  return ret;
}

// ---- flowsnake / Eisenstein (integer only) ---------------------------------------------

__device__ __forceinline__ void wb_to_flowsnake(int n,int &ex,int &ey)
// toFlowsnake(): flowsnake.cpp:92-136 (see oracle/wb_oracle.c for the derivation)
{
  int dig[11],ori=0;
  long long v=(long long)n+1235829214LL;
  #pragma unroll
  for (int i=0;i<11;i++)
  {
    dig[i]=(int)(v%7);
    v/=7;
  }
  #pragma unroll
  for (int i=10;i>=0;i--)
  {
    int t=__ldg(&g_fwdTable[ori*8+dig[i]]);
    ori=t>>4;
    dig[i]=t&7;
  }
  int x=0,y=0,px=1,py=0;
  #pragma unroll
  for (int i=0;i<11;i++)
  {
    int d=dig[i]-3,dy=(d+4)/3-1,dx=d-2*dy;
    x+=dx*px-dy*py;
    y+=dx*py+dy*px-dy*py;
    int nx=2*px+py,ny=3*py-px;
    px=nx; py=ny;
  }
  ex=x;
  ey=y;
}

__device__ __forceinline__ bool wb_from_flowsnake(int ex,int ey,long long &n)
// inverse of toFlowsnake; false if the address needs more than 11 digits
{
  int dig[11],k=0,ori=0;
  #pragma unroll
  for (int i=0;i<11;i++)
    dig[i]=3;
  while (ex || ey)
  {
    if (k>=11)
      return false;
    int d=(((ex+2*ey)%7)+10)%7-3;
    int dy=(d+4)/3-1,dx=d-2*dy;
    int a=ex-dx,b=ey-dy;
    ex=(3*a-b)/7;
    ey=(a+2*b)/7;
    dig[k++]=d+3;
  }
  long long v=0;
  #pragma unroll
  for (int i=10;i>=0;i--)
  {
    const int t=__ldg(&g_fwdTable[48+ori*8+dig[i]]);   // the inverse table: which input digit gives dig[i] in this orientation
    ori=t>>4;
    v=v*7+(t&7);
  }
  n=v-1235829214LL;
  return true;
}

__device__ __forceinline__ void wb_tile_center(int ex,int ey,const WbSnake &s,double &x,double &y)
// Flowsnake::cyl: flowsnake.cpp:263-271; Eisenstein -> complex: eisenstein.h:91-94
{
  double re=__dsub_rn((double)ex,__ddiv_rn((double)ey,2.0));
  double im=__dmul_rn((double)ey,WB_SQRT_3_4);
  re=__dmul_rn(re,s.spacing);
  im=__dmul_rn(im,s.spacing);
  x=__dadd_rn(re,s.ccx);
  y=__dadd_rn(im,s.ccy);
}

__device__ __forceinline__ int wb_covering_tiles(const WbSnake &s,double px,double py,int *nrel)
// All tiles whose cylinder (Cylinder::in, shape.cpp:214-218: hypot <= radius) contains the
// point; candidates are the 19 lattice addresses within hex distance 2 of the rounded one (a
// superset of every centre within 41/71 spacing however the rounding falls).  Two stages, so that the
// lanes of a warp stay together: first the cheap squared-distance test over all 19 (the few survivors
// are remembered), then libm's hypot and the inverse flowsnake for the survivors only — in one loop
// the long inverse ran in whichever of the 19 iterations some lane happened to hit (2.0x slower).
// Returns the count (<= 3 in practice, capped at 4) and the sequence numbers minus lo.
{
  double u=(px-s.ccx)/s.spacing,v=(py-s.ccy)/s.spacing;
  int y0=(int)wb_lrint(v/WB_SQRT_3_4),x0=(int)wb_lrint(u+y0*0.5),cnt=0;
  double r2hi=s.radius*s.radius*(1+1e-9);
  uint32_t hits=0;                                   // survivors of the first stage: bit (dy+2)*5+(dx+2)
  for (int dy=-2;dy<=2;dy++)
    for (int dx=-2;dx<=2;dx++)
    {
      if (dx-dy>2 || dy-dx>2)
        continue;
      double cx,cy;
      wb_tile_center(x0+dx,y0+dy,s,cx,cy);
      double ddx=__dsub_rn(cx,px),ddy=__dsub_rn(cy,py);
      if (ddx*ddx+ddy*ddy>r2hi)
        continue;
      hits|=1u<<((dy+2)*5+(dx+2));
    }
  while (hits)                                       // ascending (dy,dx), as the one-loop form visited them
  {
    const int code=__ffs(hits)-1,dy=code/5-2,dx=code%5-2;
    hits&=hits-1;
    int ex=x0+dx,ey=y0+dy;
    double cx,cy;
    wb_tile_center(ex,ey,s,cx,cy);
    double ddx=__dsub_rn(cx,px),ddy=__dsub_rn(cy,py);
    if (!(wb_hypot(ddx,ddy)<=s.radius))
      continue;
    long long n;
    if (!wb_from_flowsnake(ex,ey,n) || n<s.lo || n>s.hi)
      continue;
    if (cnt<4)
      nrel[cnt++]=(int)(n-s.lo);
  }
  return cnt;
}
