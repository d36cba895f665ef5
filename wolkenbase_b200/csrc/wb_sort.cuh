// wb_sort.cuh — device primitives: exclusive scan (u32) and a stable LSD radix sort of
// (u64 key, u32 value) pairs, 8 bits per pass.
//
// Sort pass = upsweep (per-block digit histogram) -> scan of the [digit][block] table ->
// downsweep (stable in-block ranking with __match_any_sync, staged through shared memory so
// that each digit's run leaves the block as contiguous, coalesced stores).
// HBM traffic per pass: read 8 B (upsweep) + read 12 B + write 12 B (downsweep) per element.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define WB_SORT_THREADS 256
#ifndef WB_SORT_ITEMS
#define WB_SORT_ITEMS 12
#endif
#ifndef WB_SORT_MINBLOCKS
#define WB_SORT_MINBLOCKS 4        // 64 registers: 4 CTAs/SM (measured 3 -> 4: 10.5 -> 8.8 ms for 8 passes over 1e8 pairs)
#endif
#define WB_SORT_TILE (WB_SORT_THREADS*WB_SORT_ITEMS)
#define WB_SORT_WARPS (WB_SORT_THREADS/32)

// ---------------------------------------------------------------- exclusive scan (u32)

#define WB_SCAN_THREADS 256
#define WB_SCAN_ITEMS 8
#define WB_SCAN_TILE (WB_SCAN_THREADS*WB_SCAN_ITEMS)

__device__ __forceinline__ uint32_t wb_warp_incl_scan(uint32_t v)
{
  const int lane=threadIdx.x&31;
  #pragma unroll
  for (int o=1;o<32;o<<=1)
  {
    uint32_t t=__shfl_up_sync(0xffffffffu,v,o);
    if (lane>=o)
      v+=t;
  }
  return v;
}

__device__ __forceinline__ uint32_t wb_block_excl_scan(uint32_t v,uint32_t *total,uint32_t *sm /* >=33 */)
// exclusive scan of one value per thread over the block; *total = block sum
{
  const int lane=threadIdx.x&31,w=threadIdx.x>>5,nw=blockDim.x>>5;
  uint32_t inc=wb_warp_incl_scan(v);
  if (lane==31)
    sm[w]=inc;
  __syncthreads();
  if (w==0)
  {
    uint32_t x=lane<nw?sm[lane]:0;
    uint32_t xi=wb_warp_incl_scan(x);
    sm[lane]=xi-x;
    if (lane==31)
      sm[32]=xi;
  }
  __syncthreads();
  uint32_t r=inc-v+sm[w];
  *total=sm[32];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(WB_SCAN_THREADS)
wb_scan_reduce_kernel(const uint32_t *__restrict__ in,uint64_t n,uint32_t *__restrict__ blockSums)
{
  __shared__ uint32_t sm[33];
  uint64_t base=(uint64_t)blockIdx.x*WB_SCAN_TILE;
  uint32_t s=0;
  #pragma unroll
  for (int i=0;i<WB_SCAN_ITEMS;i++)
  {
    uint64_t j=base+(uint64_t)i*WB_SCAN_THREADS+threadIdx.x;
    if (j<n)
      s+=in[j];
  }
  uint32_t tot;
  wb_block_excl_scan(s,&tot,sm);
  if (threadIdx.x==0)
    blockSums[blockIdx.x]=tot;
}

__global__ void __launch_bounds__(1024)
wb_scan_single_kernel(uint32_t *data,uint32_t n,uint32_t *total)
// in-place exclusive scan of a small array by ONE block
{
  __shared__ uint32_t sm[33];
  __shared__ uint32_t carry;
  if (threadIdx.x==0)
    carry=0;
  __syncthreads();
  for (uint32_t base=0;base<n;base+=blockDim.x)
  {
    uint32_t j=base+threadIdx.x;
    uint32_t v=j<n?data[j]:0,tot;
    uint32_t e=wb_block_excl_scan(v,&tot,sm);
    if (j<n)
      data[j]=e+carry;
    __syncthreads();
    if (threadIdx.x==0)
      carry+=tot;
    __syncthreads();
  }
  if (total && threadIdx.x==0)
    *total=carry;
}

__global__ void __launch_bounds__(WB_SCAN_THREADS)
wb_scan_apply_kernel(const uint32_t *__restrict__ in,uint32_t *__restrict__ out,uint64_t n,
                     const uint32_t *__restrict__ blockOffsets)
// out[j] = exclusive prefix of in; each thread owns WB_SCAN_ITEMS consecutive elements
{
  __shared__ uint32_t sm[33];
  uint64_t base=(uint64_t)blockIdx.x*WB_SCAN_TILE+(uint64_t)threadIdx.x*WB_SCAN_ITEMS;
  uint32_t v[WB_SCAN_ITEMS],s=0;
  #pragma unroll
  for (int i=0;i<WB_SCAN_ITEMS;i++)
  {
    v[i]=base+i<n?in[base+i]:0;
    s+=v[i];
  }
  uint32_t tot;
  uint32_t e=wb_block_excl_scan(s,&tot,sm)+blockOffsets[blockIdx.x];
  #pragma unroll
  for (int i=0;i<WB_SCAN_ITEMS;i++)
  {
    if (base+i<n)
      out[base+i]=e;
    e+=v[i];
  }
}

// ---------------------------------------------------------------- radix sort

template <typename K>                  // K = uint64_t (Morton / Hilbert keys) or uint32_t (tile numbers of the membership pairs)
__global__ void __launch_bounds__(WB_SORT_THREADS)
wb_sort_upsweep_kernel(const K *__restrict__ keys,uint64_t n,int shift,
                       uint32_t *__restrict__ table,uint32_t nBlocks)
// table[digit*nBlocks+block] = number of keys of this block's tile with that digit
{
  __shared__ uint32_t hist[256];
  hist[threadIdx.x]=0;
  __syncthreads();
  uint64_t base=(uint64_t)blockIdx.x*WB_SORT_TILE;
  #pragma unroll 4
  for (int i=0;i<WB_SORT_ITEMS;i++)
  {
    uint64_t j=base+(uint64_t)i*WB_SORT_THREADS+threadIdx.x;
    if (j<n)
      atomicAdd(&hist[(keys[j]>>shift)&255],1u);
  }
  __syncthreads();
  table[(uint64_t)threadIdx.x*nBlocks+blockIdx.x]=hist[threadIdx.x];
}

template <typename K>
__global__ void __launch_bounds__(WB_SORT_THREADS,WB_SORT_MINBLOCKS)
wb_sort_downsweep_kernel(const K *__restrict__ keysIn,const uint32_t *__restrict__ valsIn,
                         K *__restrict__ keysOut,uint32_t *__restrict__ valsOut,
                         uint64_t n,int shift,const uint32_t *__restrict__ table,uint32_t nBlocks)
{
  // warp w owns elements [w*32*ITEMS,(w+1)*32*ITEMS) of the tile, visited round by round
  // (round r, lane l -> element r*32+l), which is their input order: ranks are stable.
  __shared__ uint32_t warpCnt[WB_SORT_WARPS][256];
  __shared__ uint32_t digitBase[256];     // start of each digit's run inside the tile (local order)
  __shared__ uint32_t globalBase[256];    // where that run goes in the output
  __shared__ K skeys[WB_SORT_TILE];
  __shared__ uint32_t svals[WB_SORT_TILE];
  __shared__ uint32_t smscan[33];
  const int lane=threadIdx.x&31,w=threadIdx.x>>5;
  const uint64_t tileBase=(uint64_t)blockIdx.x*WB_SORT_TILE;
  const uint64_t warpBase=tileBase+(uint64_t)w*32*WB_SORT_ITEMS;
  for (int d=lane;d<256;d+=32)
    warpCnt[w][d]=0;
  __syncwarp();
  K key[WB_SORT_ITEMS];
  uint16_t off[WB_SORT_ITEMS];
  #pragma unroll
  for (int r=0;r<WB_SORT_ITEMS;r++)
  {
    uint64_t j=warpBase+(uint64_t)r*32+lane;
    key[r]=j<n?keysIn[j]:(K)~(K)0;
  }
  #pragma unroll
  for (int r=0;r<WB_SORT_ITEMS;r++)
  {
    uint64_t j=warpBase+(uint64_t)r*32+lane;
    bool ok=j<n;
    uint32_t d=ok?(uint32_t)((key[r]>>shift)&255):256u;   // 256 = padding, never ranked
    uint32_t peers=__match_any_sync(0xffffffffu,d);
    uint32_t rank=__popc(peers&((1u<<lane)-1));
    int leader=__ffs(peers)-1;
    uint32_t base=0;
    if (ok && lane==leader)
    {
      base=warpCnt[w][d];
      warpCnt[w][d]=base+__popc(peers);
    }
    base=__shfl_sync(0xffffffffu,base,leader);
    off[r]=(uint16_t)(base+rank);
    __syncwarp();
  }
  __syncthreads();
  // per digit: exclusive prefix over warps, digit totals, then exclusive scan over digits
  {
    const int d=threadIdx.x;
    uint32_t s=0;
    #pragma unroll
    for (int ww=0;ww<WB_SORT_WARPS;ww++)
    {
      uint32_t c=warpCnt[ww][d];
      warpCnt[ww][d]=s;
      s+=c;
    }
    uint32_t tot;
    uint32_t e=wb_block_excl_scan(s,&tot,smscan);
    digitBase[d]=e;
    globalBase[d]=table[(uint64_t)d*nBlocks+blockIdx.x];
  }
  __syncthreads();
  // place every element at its local sorted position in shared memory
  #pragma unroll
  for (int r=0;r<WB_SORT_ITEMS;r++)
  {
    uint64_t j=warpBase+(uint64_t)r*32+lane;
    if (j<n)
    {
      uint32_t d=(uint32_t)((key[r]>>shift)&255);
      uint32_t p=digitBase[d]+warpCnt[w][d]+off[r];
      skeys[p]=key[r];
      svals[p]=valsIn[j];                           // values are only needed now: keeps 12 registers free during ranking
    }
  }
  __syncthreads();
  // coalesced write-out: consecutive local positions of one digit are consecutive in the output
  uint32_t cnt=(uint32_t)(n-tileBase<WB_SORT_TILE?n-tileBase:WB_SORT_TILE);
  for (uint32_t p=threadIdx.x;p<cnt;p+=WB_SORT_THREADS)
  {
    K k=skeys[p];
    uint32_t d=(uint32_t)((k>>shift)&255);
    uint64_t dst=(uint64_t)globalBase[d]+(p-digitBase[d]);
    keysOut[dst]=k;
    valsOut[dst]=svals[p];
  }
}

// Host-side drivers ------------------------------------------------------------------------

struct WbScratch
{
  uint32_t *table=nullptr;      // radix table / scan block sums
  uint64_t tableCap=0;
  uint32_t *blockSums=nullptr;
  uint64_t blockSumsCap=0;
};

static inline uint64_t wb_div_up(uint64_t a,uint64_t b) { return (a+b-1)/b; }

#ifndef WB_NO_HOST_LAUNCH    // the SIMT emulator (tests/simt) compiles the kernels only and replays these launches itself
// exclusive scan of n u32 (n < 2^32 * tile), out may alias in.  Launch count returned via *launches.
static cudaError_t wb_exclusive_scan(const uint32_t *in,uint32_t *out,uint64_t n,uint32_t *blockSums,
                                     uint64_t blockSumsCap,cudaStream_t st,uint64_t *launches)
{
  if (n==0)
    return cudaSuccess;
  uint64_t nb=wb_div_up(n,WB_SCAN_TILE);
  if (nb>blockSumsCap)
    return cudaErrorInvalidValue;
  wb_scan_reduce_kernel<<<(unsigned)nb,WB_SCAN_THREADS,0,st>>>(in,n,blockSums);
  wb_scan_single_kernel<<<1,1024,0,st>>>(blockSums,(uint32_t)nb,nullptr);
  wb_scan_apply_kernel<<<(unsigned)nb,WB_SCAN_THREADS,0,st>>>(in,out,n,blockSums);
  if (launches)
    *launches+=3;
  return cudaGetLastError();
}

// Stable sort of (key,val) by key bits [beginBit,endBit).  Result ends in (keysA,valsA) if the
// number of passes is even, else in (keysB,valsB); *inA tells which.
template <typename K>
static cudaError_t wb_radix_sort(K *keysA,uint32_t *valsA,K *keysB,uint32_t *valsB,
                                 uint64_t n,int beginBit,int endBit,uint32_t *table,uint64_t tableCap,
                                 uint32_t *blockSums,uint64_t blockSumsCap,cudaStream_t st,
                                 bool *inA,uint64_t *launches)
{
  *inA=true;
  if (n==0)
    return cudaSuccess;
  uint64_t nb=wb_div_up(n,WB_SORT_TILE);
  if (nb*256>tableCap)
    return cudaErrorInvalidValue;
  K *ki=keysA,*ko=keysB;
  uint32_t *vi=valsA,*vo=valsB;
  for (int shift=beginBit;shift<endBit;shift+=8)
  {
    wb_sort_upsweep_kernel<K><<<(unsigned)nb,WB_SORT_THREADS,0,st>>>(ki,n,shift,table,(uint32_t)nb);
    if (launches)
      (*launches)++;
    cudaError_t e=wb_exclusive_scan(table,table,nb*256,blockSums,blockSumsCap,st,launches);
    if (e!=cudaSuccess)
      return e;
    wb_sort_downsweep_kernel<K><<<(unsigned)nb,WB_SORT_THREADS,0,st>>>(ki,vi,ko,vo,n,shift,table,(uint32_t)nb);
    if (launches)
      (*launches)++;
    K *tk=ki; ki=ko; ko=tk;
    uint32_t *tv=vi; vi=vo; vo=tv;
    *inA=!*inA;
  }
  return cudaGetLastError();
}
#endif
