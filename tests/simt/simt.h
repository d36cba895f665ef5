// simt.h — one CUDA block as a set of fibers (one per thread) in lock step: just enough of the SIMT execution model
// to run the kernels of wolkenbase_b200/csrc on a CPU, for tests.  A warp-wide *_sync intrinsic parks its lane until the
// warp's live lanes have all arrived (at the same source line, or the emulator aborts), __syncthreads until the
// block's have; blocks run one after another.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#if defined(__x86_64__)
#define SIMT_ASM_SWITCH 1
#else
#include <ucontext.h>
#endif

namespace simt
{
#ifdef SIMT_ASM_SWITCH
// A fiber switch that saves what the System V ABI makes the callee keep (rbx, rbp, r12-r15, the stack pointer, the
// x87 and SSE control words) and nothing else: ~20x cheaper than swapcontext, which also makes a signal-mask syscall.
struct ucontext_t { void *sp; };
extern "C" void simt_switch(void **from_sp,void *to_sp);
asm(R"(
.text
.globl simt_switch
.type simt_switch,@function
simt_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  subq $8,%rsp
  stmxcsr (%rsp)
  fnstcw 4(%rsp)
  movq %rsp,(%rdi)
  movq %rsi,%rsp
  ldmxcsr (%rsp)
  fldcw 4(%rsp)
  addq $8,%rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size simt_switch,.-simt_switch
)");
inline void swapcontext(ucontext_t *from,ucontext_t *to) { simt_switch(&from->sp,to->sp); }
inline void make_fiber(ucontext_t *c,char *stack,size_t bytes,void (*entry)())
{
  // initial frame: [mxcsr|fpucw] r15 r14 r13 r12 rbx rbp <entry> <pad>; after `ret` rsp is 16n+8 as at a call
  uintptr_t top=((uintptr_t)(stack+bytes))&~(uintptr_t)15;
  void **sp=(void **)top;
  *--sp=nullptr;                   // pad: keeps (rsp+8)%16==0 inside entry
  *--sp=(void *)entry;             // return address
  for (int i=0;i<6;i++)
    *--sp=nullptr;                 // rbp rbx r12 r13 r14 r15
  unsigned cw[2];
  asm volatile("stmxcsr %0" : "=m"(cw[0]));
  unsigned short fcw;
  asm volatile("fnstcw %0" : "=m"(fcw));
  cw[1]=fcw;
  *--sp=nullptr;
  memcpy(sp,cw,8);
  c->sp=sp;
}
#else
inline void make_fiber(ucontext_t *c,char *stack,size_t bytes,void (*entry)())
{
  getcontext(c);
  c->uc_stack.ss_sp=stack;
  c->uc_stack.ss_size=bytes;
  c->uc_link=nullptr;
  makecontext(c,entry,0);
}
#endif

struct dim3_ { unsigned x=1,y=1,z=1; };
enum Op { OP_NONE=0,OP_SYNC,OP_BALLOT,OP_OR,OP_MINU,OP_MAXU,OP_SHFL,OP_MATCH,OP_BARRIER };

struct Lane
{
  dim3_ tid,bid,bdim,gdim;
  ucontext_t ctx;
  char *stack=nullptr;
  int state=2;                       // 0 runnable, 1 parked at a collective, 2 finished
  // pending collective
  int op=OP_NONE,site=0;
  unsigned long long val=0,result=0;
  unsigned aux=0,mask=0;
};

struct Block
{
  Lane lane[1024];
  int nLanes=0;
  ucontext_t sched;
  int current=-1;
  std::function<void()> body;
  unsigned long long collectives=0;
};
typedef Block Warp;

inline Block *&warp_ptr() { static thread_local Block *w=nullptr; return w; }
// loop-trip counters the kernels bump through WB_EMU_COUNT(slot), once per warp and trip (lane 0 counts)
inline unsigned long long *emu_counts() { static thread_local unsigned long long c[16]; return c; }
// how often each call site (source line) of a warp-wide intrinsic was reached: a poor man's profile of the kernel
inline unsigned long long *site_counts() { static thread_local unsigned long long c[4096]; return c; }
inline Lane *cur() { Block *w=warp_ptr(); return &w->lane[w->current]; }

inline unsigned long long collective(int op,int site,unsigned long long val,unsigned aux,unsigned mask)
// Called by a lane: park until every live lane of its warp (OP_BARRIER: of its block) has arrived, then return
// this lane's result.
{
  Block *w=warp_ptr();
  Lane *l=&w->lane[w->current];
  l->op=op; l->site=site; l->val=val; l->aux=aux; l->mask=mask;
  l->state=1;
  swapcontext(&l->ctx,&w->sched);
  return l->result;
}

inline void lane_entry()
{
  Block *w=warp_ptr();
  w->body();
  Lane &l=w->lane[w->current];
  l.state=2;
  l.op=OP_NONE;
  swapcontext(&l.ctx,&w->sched);
}

inline bool resolve_warp(Block *w,int base)
// lanes base..base+31: if every live lane is parked at a warp-wide intrinsic, compute the results and make them
// runnable.  Returns false when the warp has nothing to resolve (all finished, or waiting at the block barrier).
{
  int op=OP_NONE,site=0,first=-1;
  const int end=base+32<w->nLanes?base+32:w->nLanes;
  for (int i=base;i<end;i++)
    if (w->lane[i].state==1)
    {
      if (first<0) { first=i; op=w->lane[i].op; site=w->lane[i].site; }
      else if (w->lane[i].op!=op || w->lane[i].site!=site)
      {
        fprintf(stderr,"simt: divergent collective: lane %d at line %d (op %d), lane %d at line %d (op %d)\n",
                first,site,op,i,w->lane[i].site,w->lane[i].op);
        abort();
      }
    }
    else if (w->lane[i].state==0)
      return false;                                   // somebody still has to run
  if (first<0 || op==OP_BARRIER)
    return false;
  w->collectives++;
  site_counts()[site&4095]++;
  unsigned ballot=0;
  unsigned long long acc_or=0,acc_min=~0ull,acc_max=0;
  for (int i=base;i<end;i++)
    if (w->lane[i].state==1)
    {
      if (w->lane[i].val) ballot|=1u<<(i-base);
      acc_or|=w->lane[i].val;
      if (w->lane[i].val<acc_min) acc_min=w->lane[i].val;
      if (w->lane[i].val>acc_max) acc_max=w->lane[i].val;
    }
  for (int i=base;i<end;i++)
  {
    Lane &l=w->lane[i];
    if (l.state!=1)
      continue;
    switch (op)
    {
      case OP_SYNC: l.result=0; break;
      case OP_BALLOT: l.result=ballot&l.mask; break;
      case OP_OR: l.result=acc_or; break;
      case OP_MINU: l.result=acc_min; break;
      case OP_MAXU: l.result=acc_max; break;
      case OP_SHFL:
      {
        const int src=base+(int)(l.aux&31);
        const bool there=src<end && w->lane[src].state==1;
        l.result=there?w->lane[src].val:l.val;        // reading an exited lane is undefined on the GPU; keep own value
        break;
      }
      case OP_MATCH:
      {
        unsigned m=0;
        for (int j=base;j<end;j++)
          if (w->lane[j].state==1 && w->lane[j].val==l.val)
            m|=1u<<(j-base);
        l.result=m&l.mask;
        break;
      }
      default: l.result=0;
    }
  }
  for (int i=base;i<end;i++)
    if (w->lane[i].state==1)
      w->lane[i].state=0;
  return true;
}

// Run `body` once per thread of one block of nThreads threads (threadIdx.x = firstThread .. firstThread+nThreads-1;
// firstThread > 0 runs one warp of a larger block on its own, for kernels whose warps do not interact).
inline unsigned long long run_block(const std::function<void()> &body,unsigned nThreads,unsigned firstThread,unsigned block,
                                    unsigned blockDim,unsigned gridDim,size_t stackBytes=128*1024)
{
  static thread_local Block *w=nullptr;
  if (!w)
    w=new Block;
  if (nThreads>1024) { fprintf(stderr,"simt: block of %u threads\n",nThreads); abort(); }
  warp_ptr()=w;
  w->body=body;
  w->collectives=0;
  w->nLanes=(int)nThreads;
  for (unsigned i=0;i<nThreads;i++)
  {
    Lane &l=w->lane[i];
    if (!l.stack)
      l.stack=(char *)malloc(stackBytes);
    l.tid.x=firstThread+i; l.bid.x=block; l.bdim.x=blockDim; l.gdim.x=gridDim;
    l.state=0;
    l.op=OP_NONE;
    make_fiber(&l.ctx,l.stack,stackBytes,lane_entry);
  }
  while (true)
  {
    bool live=false,progressed=false;
    for (unsigned i=0;i<nThreads;i++)
      if (w->lane[i].state==0)
      {
        w->current=(int)i;
        swapcontext(&w->sched,&w->lane[i].ctx);     // runs until the lane parks at a collective or finishes
        progressed=true;
      }
    for (unsigned base=0;base<nThreads;base+=32)
      if (resolve_warp(w,(int)base))
        progressed=true;
    // block barrier: released when every live lane waits at one
    bool allAtBarrier=true;
    for (unsigned i=0;i<nThreads;i++)
      if (w->lane[i].state!=2)
      {
        live=true;
        if (!(w->lane[i].state==1 && w->lane[i].op==OP_BARRIER))
          allAtBarrier=false;
      }
    if (!live)
      break;
    if (allAtBarrier)
    {
      for (unsigned i=0;i<nThreads;i++)
        if (w->lane[i].state==1)
        {
          w->lane[i].result=0;
          w->lane[i].state=0;
        }
      w->collectives++;
      progressed=true;
    }
    if (!progressed)
    {
      fprintf(stderr,"simt: deadlock in block %u (a warp split between __syncthreads and a warp-wide intrinsic?)\n",block);
      abort();
    }
  }
  return w->collectives;
}

inline unsigned long long run_warp(const std::function<void()> &body,unsigned firstThread,unsigned block,unsigned blockDim,
                                   unsigned gridDim)
{
  return run_block(body,32,firstThread,block,blockDim,gridDim);
}
} // namespace simt
