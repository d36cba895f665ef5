// cuda_runtime.h — stand-in that lets the device code of wolkenbase_b200/csrc compile as host C++ and run
// under the SIMT emulator of tests/simt/simt.h.  TEST INFRASTRUCTURE ONLY (never on the product path: the
// product is the nvcc build of the same sources).
//
// What is emulated: one warp = 32 fibers (ucontext) run in lock step between warp-wide intrinsics; every
// *_sync intrinsic is a rendezvous of the warp's live lanes (the emulator checks that all lanes arrive at the
// same call site, i.e. that the code is warp-uniform where CUDA requires it).  Arithmetic intrinsics map to
// the IEEE operation they name (compile with -ffp-contract=off); __fdividef is an exact float division, so
// masks built from approximate angles may differ from the GPU's by a rounding — never results.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <climits>
#include <algorithm>
#include "../simt.h"

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static thread_local
#define __constant__

using std::min;
using std::max;
using std::isnan;
using std::isinf;

struct uint4 { unsigned x,y,z,w; };
inline uint4 make_uint4(unsigned x,unsigned y,unsigned z,unsigned w) { uint4 r={x,y,z,w}; return r; }
struct simt_dim3 { unsigned x=1,y=1,z=1; };
#define threadIdx (simt::cur()->tid)
#define blockIdx (simt::cur()->bid)
#define blockDim (simt::cur()->bdim)
#define gridDim (simt::cur()->gdim)

#define WB_EMU_COUNT(slot) do { if ((simt::cur()->tid.x&31)==0) simt::emu_counts()[(slot)&15]++; } while (0)

// ---- arithmetic --------------------------------------------------------------------------------------
inline double __dmul_rn(double a,double b) { return a*b; }
inline double __dadd_rn(double a,double b) { return a+b; }
inline double __dsub_rn(double a,double b) { return a-b; }
inline double __ddiv_rn(double a,double b) { return a/b; }
inline double __fma_rn(double a,double b,double c) { return std::fma(a,b,c); }
inline long long __double2ll_rn(double v) { return std::llrint(v); }
inline float __fdividef(float a,float b) { return a/b; }
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u,&f,4); return u; }
inline float __uint_as_float(unsigned u) { float f; memcpy(&f,&u,4); return f; }
inline long long __double_as_longlong(double d) { long long l; memcpy(&l,&d,8); return l; }
inline double __longlong_as_double(long long l) { double d; memcpy(&d,&l,8); return d; }
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline unsigned __funnelshift_r(unsigned lo,unsigned hi,unsigned s)
{
  unsigned long long v=((unsigned long long)hi<<32)|lo;
  return (unsigned)(v>>(s&31));
}
template <typename T> inline T __ldg(const T *p) { return *p; }

// ---- memory (one warp runs at a time on this OS thread: plain operations) -------------------------------
template <typename T,typename U> inline T atomicAdd(T *p,U v) { T o=*p; *p=(T)(o+(T)v); return o; }
template <typename T,typename U> inline T atomicMax(T *p,U v) { T o=*p; if ((T)v>o) *p=(T)v; return o; }
template <typename T,typename U> inline T atomicMin(T *p,U v) { T o=*p; if ((T)v<o) *p=(T)v; return o; }

// ---- warp-wide intrinsics -----------------------------------------------------------------------------
#define SIMT_SITE (__builtin_LINE())
inline void __syncwarp(unsigned mask=0xffffffffu,int site=SIMT_SITE) { simt::collective(simt::OP_SYNC,site,0,0,mask); }
inline void __syncthreads(int site=SIMT_SITE) { simt::collective(simt::OP_BARRIER,site,0,0,0xffffffffu); }
inline unsigned __ballot_sync(unsigned mask,int pred,int site=SIMT_SITE)
{ return (unsigned)simt::collective(simt::OP_BALLOT,site,pred?1:0,0,mask); }
inline int __any_sync(unsigned mask,int pred,int site=SIMT_SITE)
{ return simt::collective(simt::OP_BALLOT,site,pred?1:0,0,mask)!=0; }
inline unsigned __reduce_or_sync(unsigned mask,unsigned v,int site=SIMT_SITE)
{ return (unsigned)simt::collective(simt::OP_OR,site,v,0,mask); }
inline unsigned __reduce_min_sync(unsigned mask,unsigned v,int site=SIMT_SITE)
{ return (unsigned)simt::collective(simt::OP_MINU,site,v,0,mask); }
inline unsigned __reduce_max_sync(unsigned mask,unsigned v,int site=SIMT_SITE)
{ return (unsigned)simt::collective(simt::OP_MAXU,site,v,0,mask); }
inline unsigned __match_any_sync(unsigned mask,unsigned long long v,int site=SIMT_SITE)
{ return (unsigned)simt::collective(simt::OP_MATCH,site,v,0,mask); }
template <typename T> inline T __shfl_sync(unsigned mask,T v,int src,int width=32,int site=SIMT_SITE)
{
  static_assert(sizeof(T)<=8,"shuffle of at most 8 bytes");
  unsigned long long bits=0;
  memcpy(&bits,&v,sizeof(T));
  bits=simt::collective(simt::OP_SHFL,site,bits,(unsigned)(src&31),mask);
  T r;
  memcpy(&r,&bits,sizeof(T));
  return r;
}
template <typename T> inline T __shfl_xor_sync(unsigned mask,T v,int lanemask,int width=32,int site=SIMT_SITE)
{
  return __shfl_sync(mask,v,(int)((simt::cur()->tid.x&31)^(unsigned)lanemask),width,site);
}
template <typename T> inline T __shfl_up_sync(unsigned mask,T v,unsigned delta,int width=32,int site=SIMT_SITE)
{
  int lane=(int)(simt::cur()->tid.x&31);
  return __shfl_sync(mask,v,lane>=(int)delta?lane-(int)delta:lane,width,site);
}
template <typename T> inline T __shfl_down_sync(unsigned mask,T v,unsigned delta,int width=32,int site=SIMT_SITE)
{
  int lane=(int)(simt::cur()->tid.x&31);
  return __shfl_sync(mask,v,lane+(int)delta<32?lane+(int)delta:lane,width,site);
}
