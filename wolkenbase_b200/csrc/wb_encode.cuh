// wb_encode.cuh — K11: the output records, made on the device.
//
// Replaces the per-point loop of CloudOutput::writeFiles (cloudoutput.cpp:187-229) over
// OctStore::getAll + LasHeader::writePoint (las.cpp:822-904): every stored point is decoded from
// its original record as LasHeader::readPoint does (las.cpp:735-820), given its new class, and
// re-encoded in the output format with XYZ re-quantised to the output scale and offset.  Which
// file a bucket's points of one class go to, and where, is decided by the host (the reference's
// greedy "least full file" rule is sequential over buckets); the kernel gets a byte offset per
// (bucket, class slot) and appends the bucket's points of that class in bucket order.
#pragma once
#include <cstdint>

struct WbRecSeg { const uint8_t *recs; int fmt,recLen; };
struct WbRecSegs { WbRecSeg s[WB_MAX_SEGMENTS]; };

struct WbOutSpec
{
  int fmt,recLen,nClasses,separate;
  double scale[3],offset[3],unit;
};

#define WB_ENC_WARPS 4
#define WB_ENC_ACC 24            // per file: min xi,yi,zi, max xi,yi,zi, 2 spare, 16 per-return counts
#define WB_ENC_SMEM_FILES 448

__device__ __forceinline__ int wb_deg_to_bin(double deg)
// degtobin -> rottobin, angle.cpp:233-251
{
  double t=__ddiv_rn(__ddiv_rn(deg,360.),2.);
  double fp=__dmul_rn(2.,__dsub_rn(t,trunc(t)));
  if (fp>=1) fp=__dsub_rn(fp,2.);
  if (fp<-1) fp=__dadd_rn(fp,2.);
  return (int)__double2ll_rn(__dmul_rn(2147483648.,fp));
}

__device__ __forceinline__ double wb_bin_to_deg(int a)
{
  return __dmul_rn(__ddiv_rn((double)a,2147483648.),360.);
}

__device__ __forceinline__ uint32_t wb_rd16(const uint8_t *p) { return (uint32_t)p[0]|((uint32_t)p[1]<<8); }
__device__ __forceinline__ void wb_wr16(uint8_t *p,uint32_t v) { p[0]=(uint8_t)v; p[1]=(uint8_t)(v>>8); }
__device__ __forceinline__ void wb_wr32(uint8_t *p,uint32_t v) { wb_wr16(p,v); wb_wr16(p+2,v>>16); }

__global__ void __launch_bounds__(256)
wb_leaf_class_counts_kernel(const WbLeafDev *__restrict__ leaves,uint32_t nLeaves,const uint8_t *__restrict__ labelSorted,
                            const uint8_t *__restrict__ lut,int nClasses,uint32_t *__restrict__ counts)
// one warp per bucket: how many of its points carry the class of slot k
{
  uint32_t leaf=(uint32_t)(((unsigned long long)blockIdx.x*blockDim.x+threadIdx.x)>>5);
  int lane=threadIdx.x&31;
  if (leaf>=nLeaves)
    return;
  const unsigned long long first=leaves[leaf].first;
  const uint32_t cnt=leaves[leaf].count;
  for (int k0=0;k0<nClasses;k0+=32)
  {
    uint32_t mine=0;                                  // lane l counts slot k0+l
    for (uint32_t o=0;o<cnt;o+=32)
    {
      int k=(o+lane<cnt)?lut[labelSorted[first+o+lane]]:255;
      for (int l=0;l<32 && k0+l<nClasses;l++)
      {
        uint32_t c=__popc(__ballot_sync(0xffffffffu,k==k0+l));
        if (lane==l)
          mine+=c;
      }
    }
    if (k0+lane<nClasses)
      counts[(unsigned long long)leaf*nClasses+k0+lane]=mine;
  }
}

__global__ void __launch_bounds__(256)
wb_attr_source_kernel(const uint32_t *__restrict__ perm,unsigned long long nv,uint32_t *__restrict__ inv,
                      uint32_t *__restrict__ attrSrc)
{
  unsigned long long j=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  if (j<nv)
  {
    inv[perm[j]]=(uint32_t)j;
    attrSrc[j]=perm[j];
  }
}

__global__ void __launch_bounds__(256)
wb_attr_last_kernel(const uint32_t *__restrict__ dupIn,const uint32_t *__restrict__ dupRep,unsigned long long nDup,
                    const uint32_t *__restrict__ inv,uint32_t *attrSrc)
// OctBuffer::put overwrites the stored point with the newcomer (octree.cpp:626-644): the attributes
// that survive are those of the last record at the location, in the place of the first
{
  unsigned long long u=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  if (u<nDup)
    atomicMax(&attrSrc[inv[dupRep[u]]],dupIn[u]);
}

__global__ void __launch_bounds__(WB_ENC_WARPS*32)
wb_encode_kernel(const WbLeafDev *__restrict__ leaves,uint32_t nLeaves,
                 const double *__restrict__ sx,const double *__restrict__ sy,const double *__restrict__ sz,
                 const uint8_t *__restrict__ labelSorted,const uint32_t *__restrict__ src,
                 const WbSegments *__restrict__ segs,const WbRecSegs *__restrict__ rsegs,
                 WbOutSpec spec,const uint8_t *__restrict__ lut,
                 const unsigned long long *__restrict__ dest,const uint32_t *__restrict__ fileOf,uint32_t nFiles,
                 uint8_t *__restrict__ out,int *gMinMax,unsigned long long *gCount)
{
  extern __shared__ int wbEncSmem[];
  const bool accSmem=nFiles<=WB_ENC_SMEM_FILES;
  const int K=spec.separate?spec.nClasses:1;
  int *acc=wbEncSmem;                                              // [nFiles][WB_ENC_ACC] when accSmem
  uint32_t *base=(uint32_t *)(wbEncSmem+(accSmem?nFiles*WB_ENC_ACC:0));   // [WB_ENC_WARPS][K]
  const int warp=threadIdx.x>>5,lane=threadIdx.x&31;
  if (accSmem)
    for (uint32_t i=threadIdx.x;i<nFiles*WB_ENC_ACC;i+=blockDim.x)
    {
      int w=i%WB_ENC_ACC;
      acc[i]=w<3?0x7fffffff:(w<6?(int)0x80000000:0);
    }
  for (int i=lane;i<K;i+=32)
    base[warp*K+i]=0;
  __syncthreads();
  const uint32_t leaf=blockIdx.x*WB_ENC_WARPS+warp;
  if (leaf<nLeaves)
  {
    const unsigned long long first=leaves[leaf].first;
    const uint32_t cnt=leaves[leaf].count;
    for (uint32_t o=0;o<cnt;o+=32)
    {
      const bool valid=o+lane<cnt;
      const unsigned long long j=first+o+lane;
      int lab=0,k=255;
      if (valid)
      {
        lab=labelSorted[j];
        k=spec.separate?lut[lab]:0;
      }
      const uint32_t peers=__match_any_sync(0xffffffffu,k);
      const uint32_t rank=__popc(peers&((1u<<lane)-1));
      uint32_t b=0;
      if (k!=255)
        b=base[warp*K+k];
      __syncwarp();
      if (k!=255 && rank==0)
        base[warp*K+k]=b+__popc(peers);
      __syncwarp();
      if (k==255)
        continue;                                   // class without an output file (cloudoutput.cpp:212-214)
      const unsigned long long slot=(unsigned long long)leaf*K+k;
      const uint32_t f=fileOf[slot];
      uint8_t *w=out+dest[slot]+(unsigned long long)(b+rank)*spec.recLen;
      // ---- LasHeader::readPoint (las.cpp:735-820)
      const uint32_t i=src[j];
      int lo=0,hi=segs->n-1;
      while (lo<hi)
      {
        int mid=(lo+hi+1)>>1;
        if (segs->s[mid].first<=i)
          lo=mid;
        else
          hi=mid-1;
      }
      const int fi=rsegs->s[lo].fmt;
      const uint8_t *r=rsegs->s[lo].recs+(unsigned long long)(i-segs->s[lo].first)*rsegs->s[lo].recLen;
      uint32_t intensity=wb_rd16(r+12),returnNum,nReturns,scanDir,edge,flags,channel=0,user,source;
      int angle,oi;
      if (fi<6)
      {
        returnNum=r[14]&7; nReturns=(r[14]>>3)&7; scanDir=(r[14]>>6)&1; edge=(r[14]>>7)&1;
        flags=(r[15]>>5)&7;
        angle=wb_deg_to_bin((double)(signed char)r[16]);
        user=r[17];
        source=wb_rd16(r+18);
        oi=20;
      }
      else
      {
        returnNum=r[14]&15; nReturns=(r[14]>>4)&15;
        flags=r[15]&15; channel=(r[15]>>4)&3; scanDir=(r[15]>>6)&1; edge=(r[15]>>7)&1;
        user=r[17];
        angle=wb_deg_to_bin(__dmul_rn((double)(short)wb_rd16(r+18),0.006));
        source=wb_rd16(r+20);
        oi=22;
      }
      unsigned long long gps=0;
      uint32_t red=0,green=0,blue=0,nir=0;
      if ((1<<fi)&0x7fa)
      {
        for (int q=7;q>=0;q--)
          gps=(gps<<8)|r[oi+q];
        oi+=8;
      }
      if ((1<<fi)&0x5ac)
      {
        red=wb_rd16(r+oi); green=wb_rd16(r+oi+2); blue=wb_rd16(r+oi+4);
        oi+=6;
      }
      if ((1<<fi)&0x500)
        nir=wb_rd16(r+oi);
      if (returnNum==0)
        returnNum=1;                                // a stored point with return number 0 comes from a
                                                    // keep-zeros file, threads.cpp:527-528
      // ---- LasHeader::writePoint (las.cpp:822-904)
      const int xi=(int)__double2ll_rn(__ddiv_rn(__dsub_rn(__ddiv_rn(sx[j],spec.unit),spec.offset[0]),spec.scale[0]));
      const int yi=(int)__double2ll_rn(__ddiv_rn(__dsub_rn(__ddiv_rn(sy[j],spec.unit),spec.offset[1]),spec.scale[1]));
      const int zi=(int)__double2ll_rn(__ddiv_rn(__dsub_rn(__ddiv_rn(sz[j],spec.unit),spec.offset[2]),spec.scale[2]));
      uint8_t rec[40];
      #pragma unroll
      for (int q=0;q<40;q++)
        rec[q]=0;
      wb_wr32(rec,(uint32_t)xi); wb_wr32(rec+4,(uint32_t)yi); wb_wr32(rec+8,(uint32_t)zi);
      wb_wr16(rec+12,intensity);
      int oo;
      if (spec.fmt<6)
      {
        rec[14]=(uint8_t)((returnNum&7)+((nReturns&7)<<3)+((scanDir&1)<<6)+((edge&1)<<7));
        rec[15]=(uint8_t)((lab&31)+((flags&7)<<5));
        rec[16]=(uint8_t)__double2ll_rn(wb_bin_to_deg(angle));
        rec[17]=(uint8_t)user;
        wb_wr16(rec+18,source);
        oo=20;
      }
      else
      {
        rec[14]=(uint8_t)((returnNum&15)+((nReturns&15)<<4));
        rec[15]=(uint8_t)((flags&15)+((channel&3)<<4)+((scanDir&1)<<6)+((edge&1)<<7));
        rec[16]=(uint8_t)lab;
        rec[17]=(uint8_t)user;
        wb_wr16(rec+18,(uint32_t)(unsigned short)(short)__double2ll_rn(__ddiv_rn(wb_bin_to_deg(angle),0.006)));
        wb_wr16(rec+20,source);
        oo=22;
      }
      if ((1<<spec.fmt)&0x7fa)
      {
        for (int q=0;q<8;q++)
          rec[oo+q]=(uint8_t)(gps>>(8*q));
        oo+=8;
      }
      if ((1<<spec.fmt)&0x5ac)
      {
        wb_wr16(rec+oo,red); wb_wr16(rec+oo+2,green); wb_wr16(rec+oo+4,blue);
        oo+=6;
      }
      if ((1<<spec.fmt)&0x500)
        wb_wr16(rec+oo,nir);
      for (int q=0;q<spec.recLen;q+=2)              // record lengths and file offsets are even
        *(unsigned short *)(w+q)=(unsigned short)(rec[q]|(rec[q+1]<<8));
      // ---- running header figures: integer extremes (wx=xi*scale+offset is monotone in xi) and per-return counts
      if (accSmem)
      {
        int *a=acc+f*WB_ENC_ACC;
        atomicMin(a,xi); atomicMin(a+1,yi); atomicMin(a+2,zi);
        atomicMax(a+3,xi); atomicMax(a+4,yi); atomicMax(a+5,zi);
        atomicAdd(a+8,1);
        if (returnNum<16)
          atomicAdd(a+8+returnNum,1);
      }
      else
      {
        int *a=gMinMax+(unsigned long long)f*6;
        atomicMin(a,xi); atomicMin(a+1,yi); atomicMin(a+2,zi);
        atomicMax(a+3,xi); atomicMax(a+4,yi); atomicMax(a+5,zi);
        atomicAdd(gCount+(unsigned long long)f*16,1ull);
        if (returnNum<16)
          atomicAdd(gCount+(unsigned long long)f*16+returnNum,1ull);
      }
    }
  }
  if (accSmem)
  {
    __syncthreads();
    for (uint32_t i=threadIdx.x;i<nFiles*WB_ENC_ACC;i+=blockDim.x)
    {
      const uint32_t f=i/WB_ENC_ACC;
      const int w=i%WB_ENC_ACC,v=acc[i];
      if (w<3)
      {
        if (v!=0x7fffffff) atomicMin(gMinMax+(unsigned long long)f*6+w,v);
      }
      else if (w<6)
      {
        if (v!=(int)0x80000000) atomicMax(gMinMax+(unsigned long long)f*6+w,v);
      }
      else if (w>=8 && v)
        atomicAdd(gCount+(unsigned long long)f*16+(w-8),(unsigned long long)v);
    }
  }
}

// ============================================================================ census (testpattern.cpp:56-123)
// censusPoints(): test data carries its point number as GPS time; after a write the reference walks every block of
// the store, sets one bit per number and reports numbers seen twice and numbers missing below the highest one.
// One thread per stored point, two passes: the highest number (which sizes the bit set), then the bits.
struct WbCensus { unsigned long long notInteger,duplicate,maxPlusOne; };

__global__ void __launch_bounds__(256)
wb_census_mark_kernel(const uint32_t *__restrict__ src,unsigned long long nv,const WbSegments *__restrict__ segs,
                      const WbRecSegs *__restrict__ rsegs,unsigned long long *__restrict__ bits /* NULL: first pass */,
                      WbCensus *__restrict__ out)
{
  unsigned long long j=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  const bool have=j<nv;                             // whole warps stay: the first pass votes
  const uint32_t i=have?src[j]:src[0];
  int lo=0,hi=segs->n-1;
  while (lo<hi)
  {
    int mid=(lo+hi+1)>>1;
    if (segs->s[mid].first<=i)
      lo=mid;
    else
      hi=mid-1;
  }
  const int fi=rsegs->s[lo].fmt;
  const uint8_t *r=rsegs->s[lo].recs+(unsigned long long)(i-segs->s[lo].first)*rsegs->s[lo].recLen;
  double t=0;                                       // formats without GPS time read as 0 (LasPoint ctor, las.cpp:128)
  if ((1<<fi)&0x7fa)                                // MASK_GPSTIME: 10-5, 4, 3, 1 (las.cpp:776)
  {
    unsigned long long u=0;
    const uint8_t *g=r+(fi<6?20:22);
    #pragma unroll
    for (int k=7;k>=0;k--)
      u=(u<<8)|g[k];
    t=__longlong_as_double((long long)u);
  }
  // n=lrint(t); n!=t || n<0 -> not test data (testpattern.cpp:68-70; n is an int there: numbers stay below 2^31)
  const bool bad=have && (!(t>=0) || t>=2147483648.0 || t!=rint(t));
  const unsigned long long nn=(bad || !have)?0ull:(unsigned long long)t;
  if (!bits)
  {
    // first pass: one atomic per warp, not per thread (they all aim at the same two words)
    const unsigned nBad=__popc(__ballot_sync(0xffffffffu,bad));
    const unsigned top=__reduce_max_sync(0xffffffffu,(bad || !have)?0u:(unsigned)nn+1u);
    if ((threadIdx.x&31)==0)
    {
      if (nBad)
        atomicAdd(&out->notInteger,(unsigned long long)nBad);
      if (top)
        atomicMax(&out->maxPlusOne,(unsigned long long)top);
    }
    return;
  }
  if (bad || !have)
    return;
  const unsigned long long m=1ull<<(nn&63);
  if (atomicOr(&bits[nn>>6],m)&m)
    atomicAdd(&out->duplicate,1ull);
}

__global__ void __launch_bounds__(256)
wb_census_missing_kernel(const unsigned long long *__restrict__ bits,unsigned long long top,
                         unsigned long long *__restrict__ nMissing,unsigned long long *__restrict__ list,unsigned long long cap)
// numbers below top = maxPoint that no stored point carries; `cap` of them (any) go into list
{
  unsigned long long w=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;
  if (w*64>=top)
    return;
  unsigned long long miss=~bits[w];
  if ((w+1)*64>top)
    miss&=(1ull<<(top-w*64))-1;
  if (!miss)
    return;
  unsigned long long at=atomicAdd(nMissing,(unsigned long long)__popcll(miss));
  while (miss && at<cap)
  {
    int b=__ffsll((long long)miss)-1;
    miss&=miss-1;
    list[at++]=w*64+b;
  }
}
