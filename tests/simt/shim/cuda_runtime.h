// cuda_runtime.h — stand-in that lets the device code of wolkenbase_b200/csrc compile as host C++ and run
// under the SIMT emulator of tests/simt/simt.h.  TEST INFRASTRUCTURE ONLY (never on the product path: the
// product is the nvcc build of the same sources).
//
// What is emulated: one block = one fiber per thread, run in lock step between collectives; every *_sync
// intrinsic is a rendezvous of the warp's live lanes (the emulator checks that all lanes arrive at the same call
// site, i.e. that the code is warp-uniform where CUDA requires it), __syncthreads one of the block; __shared__ is
// a per-block static; atomics are plain (blocks and lanes never run concurrently).  Arithmetic intrinsics map to
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
#define __align__(n) alignas(n)
#define __constant__

using std::min;
using std::max;
using std::isnan;
using std::isinf;

struct uint4 { unsigned x,y,z,w; };
inline uint4 make_uint4(unsigned x,unsigned y,unsigned z,unsigned w) { uint4 r={x,y,z,w}; return r; }
struct int4 { int x,y,z,w; };
inline int4 make_int4(int x,int y,int z,int w) { int4 r={x,y,z,w}; return r; }
struct float4 { float x,y,z,w; };
inline float4 make_float4(float x,float y,float z,float w) { float4 r={x,y,z,w}; return r; }
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
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline unsigned __funnelshift_r(unsigned lo,unsigned hi,unsigned s)
{
  unsigned long long v=((unsigned long long)hi<<32)|lo;
  return (unsigned)(v>>(s&31));
}
template <typename T> inline T __ldg(const T *p) { return *p; }

// ---- memory (one warp runs at a time on this OS thread: plain operations) -------------------------------
template <typename T,typename U> inline T atomicAdd(T *p,U v) { T o=*p; *p=(T)(o+(T)v); return o; }
template <typename T,typename U> inline T atomicMax(T *p,U v) { T o=*p; if ((T)v>o) *p=(T)v; return o; }
template <typename T,typename U> inline T atomicOr(T *p,U v) { T o=*p; *p=(T)(o|(T)v); return o; }
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

// ---- just enough of the CUDA runtime API for wolken_b200.cu to compile and run as host code ----------------------
// "Device" memory is the heap, streams are immediate (every call completes before it returns), events are wall-clock
// stamps.  Used by the emulated build of the whole library (tests/simt/gen, libwolken_b200_emulated.so).
#include <chrono>
#include <cstdlib>
#include <vector>
typedef int cudaError_t;
enum { cudaSuccess=0,cudaErrorInvalidValue=1,cudaErrorMemoryAllocation=2 };
enum cudaMemcpyKind { cudaMemcpyHostToHost=0,cudaMemcpyHostToDevice=1,cudaMemcpyDeviceToHost=2,cudaMemcpyDeviceToDevice=3,cudaMemcpyDefault=4 };
struct simt_stream_t { int id; };
struct simt_event_t { double t; };
typedef simt_stream_t *cudaStream_t;
typedef simt_event_t *cudaEvent_t;
enum { cudaStreamNonBlocking=1,cudaEventDisableTiming=2,cudaHostAllocDefault=0 };
inline const char *cudaGetErrorString(cudaError_t e) { return e==cudaSuccess?"no error":(e==cudaErrorMemoryAllocation?"out of memory":"invalid value"); }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n=1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
template <typename T> inline cudaError_t cudaMalloc(T **p,size_t n) { *p=(T *)malloc(n?n:1); return *p?cudaSuccess:cudaErrorMemoryAllocation; }
inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
template <typename T> inline cudaError_t cudaHostAlloc(T **p,size_t n,unsigned) { *p=(T *)malloc(n?n:1); return *p?cudaSuccess:cudaErrorMemoryAllocation; }
inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d,const void *s,size_t n,cudaMemcpyKind) { memmove(d,s,n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d,const void *s,size_t n,cudaMemcpyKind,cudaStream_t=nullptr) { memmove(d,s,n); return cudaSuccess; }
inline cudaError_t cudaMemset(void *d,int v,size_t n) { memset(d,v,n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *d,int v,size_t n,cudaStream_t=nullptr) { memset(d,v,n); return cudaSuccess; }
template <typename T> inline cudaError_t cudaMemcpyToSymbol(T &sym,const void *s,size_t n) { memcpy(&sym,s,n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s,unsigned) { *s=new simt_stream_t{0}; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s,unsigned,int) { *s=new simt_stream_t{0}; return cudaSuccess; }
inline cudaError_t cudaDeviceGetStreamPriorityRange(int *least,int *greatest) { *least=0; *greatest=-5; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t,cudaEvent_t,unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e=new simt_event_t{0}; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e,unsigned) { *e=new simt_event_t{0}; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e,cudaStream_t=nullptr)
{ e->t=std::chrono::duration<double,std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms,cudaEvent_t a,cudaEvent_t b) { *ms=(float)(b->t-a->t); return cudaSuccess; }

// kernel<<<grid,block,smem,...>>>(args) after tests/simt/gen_host_build.py: the blocks one after another, each with
// one fiber per thread; `extern __shared__` arrays point into a per-launch buffer of `smem` bytes
inline std::vector<unsigned long long> &simt_dynamic_buffer() { static thread_local std::vector<unsigned long long> b; return b; }
inline void *simt_dynamic_smem() { return simt_dynamic_buffer().data(); }
template <typename F> inline void simt_launch(unsigned long long grid,unsigned block,size_t smem,F body)
{
  simt_dynamic_buffer().assign(smem/8+2,0);
  // SIMT_BLOCK_ORDER=reverse|shuffle: the GPU runs blocks in no particular order; results must not depend on it
  static const char *order=getenv("SIMT_BLOCK_ORDER");
  if (order && order[0]=='r')
    for (unsigned long long b=grid;b-->0;)
      simt::run_block(body,block,0,(unsigned)b,block,(unsigned)grid);
  else if (order && order[0]=='s')
  {
    // a fixed odd stride visits every block once (grid and stride coprime when the stride is a large prime)
    const unsigned long long stride=2654435761ull%(grid?grid:1)|1ull;
    unsigned long long g=grid,a=stride,t;
    while (a) { t=g%a; g=a; a=t; }                     // gcd
    if (g!=1)
      for (unsigned long long b=grid;b-->0;)
        simt::run_block(body,block,0,(unsigned)b,block,(unsigned)grid);
    else
      for (unsigned long long k=0,b=grid/2;k<grid;k++,b=(b+stride)%grid)
        simt::run_block(body,block,0,(unsigned)b,block,(unsigned)grid);
  }
  else
    for (unsigned long long b=0;b<grid;b++)
      simt::run_block(body,block,0,(unsigned)b,block,(unsigned)grid);
}
