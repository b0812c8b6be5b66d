// IMAD flavour microbenchmark (sm_100a): thread-level ops/clk/SM for several multiply-add forms.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint32_t u32; typedef uint64_t u64;
#define REP16(x) x x x x x x x x x x x x x x x x
template<int MODE> __global__ void k(u32* sink, int iters){
  u32 t = blockIdx.x*blockDim.x+threadIdx.x;
  u32 x = t|1, y = (t*2654435761u)|1;
  u32 r[32];
  #pragma unroll
  for(int i=0;i<32;i++) r[i]=t+i;
  u64 a0=t,a1=t+1,a2=t+2,a3=t+3,a4=t+4,a5=t+5,a6=t+6,a7=t+7;
  for(int it=0; it<iters; ++it){
    if (MODE==0){ // mad.wide no carry, 8 chains x16
      #pragma unroll
      for(int k2=0;k2<16;k2++){
        asm volatile("mad.wide.u32 %0,%1,%2,%0;":"+l"(a0):"r"((u32)a1),"r"(y)); asm volatile("mad.wide.u32 %0,%1,%2,%0;":"+l"(a1):"r"((u32)a2),"r"(y));
        asm volatile("mad.wide.u32 %0,%1,%2,%0;":"+l"(a2):"r"((u32)a3),"r"(y)); asm volatile("mad.wide.u32 %0,%1,%2,%0;":"+l"(a3):"r"((u32)a4),"r"(y));
        asm volatile("mad.wide.u32 %0,%1,%2,%0;":"+l"(a4):"r"((u32)a5),"r"(y)); asm volatile("mad.wide.u32 %0,%1,%2,%0;":"+l"(a5):"r"((u32)a6),"r"(y));
        asm volatile("mad.wide.u32 %0,%1,%2,%0;":"+l"(a6):"r"((u32)a7),"r"(y)); asm volatile("mad.wide.u32 %0,%1,%2,%0;":"+l"(a7):"r"((u32)a0),"r"(y));
      }
    } else if (MODE==1){ // carry chains: 4 independent chains of 16 pairs (=64 fused IMAD.WIDE.X) x2 = 128
      #pragma unroll
      for(int rep=0;rep<2;rep++){
      #pragma unroll
      for(int c=0;c<4;c++){
        u32* q=&r[8*c];
        asm volatile("mad.lo.cc.u32 %0,%8,%9,%0; madc.hi.cc.u32 %1,%8,%9,%1; madc.lo.cc.u32 %2,%8,%9,%2; madc.hi.cc.u32 %3,%8,%9,%3;"
            "madc.lo.cc.u32 %4,%8,%9,%4; madc.hi.cc.u32 %5,%8,%9,%5; madc.lo.cc.u32 %6,%8,%9,%6; madc.hi.cc.u32 %7,%8,%9,%7;"
            "madc.lo.cc.u32 %0,%8,%9,%0; madc.hi.cc.u32 %1,%8,%9,%1; madc.lo.cc.u32 %2,%8,%9,%2; madc.hi.cc.u32 %3,%8,%9,%3;"
            "madc.lo.cc.u32 %4,%8,%9,%4; madc.hi.cc.u32 %5,%8,%9,%5; madc.lo.cc.u32 %6,%8,%9,%6; madc.hi.cc.u32 %7,%8,%9,%7;"
            "madc.lo.cc.u32 %0,%8,%9,%0; madc.hi.cc.u32 %1,%8,%9,%1; madc.lo.cc.u32 %2,%8,%9,%2; madc.hi.cc.u32 %3,%8,%9,%3;"
            "madc.lo.cc.u32 %4,%8,%9,%4; madc.hi.cc.u32 %5,%8,%9,%5; madc.lo.cc.u32 %6,%8,%9,%6; madc.hi.cc.u32 %7,%8,%9,%7;"
            "madc.lo.cc.u32 %0,%8,%9,%0; madc.hi.cc.u32 %1,%8,%9,%1; madc.lo.cc.u32 %2,%8,%9,%2; madc.hi.cc.u32 %3,%8,%9,%3;"
            "madc.lo.cc.u32 %4,%8,%9,%4; madc.hi.cc.u32 %5,%8,%9,%5; madc.lo.cc.u32 %6,%8,%9,%6; madc.hi.u32 %7,%8,%9,%7;"
            :"+r"(q[0]),"+r"(q[1]),"+r"(q[2]),"+r"(q[3]),"+r"(q[4]),"+r"(q[5]),"+r"(q[6]),"+r"(q[7]):"r"(r[(8*c+9)&31]),"r"(y));
      }}
    } else if (MODE==2){ // mad.lo 32-bit only, 32 chains x4 = 128
      #pragma unroll
      for(int k2=0;k2<4;k2++){
        #pragma unroll
        for(int i=0;i<32;i++) asm volatile("mad.lo.u32 %0,%1,%2,%0;":"+r"(r[i]):"r"(r[(i+1)&31]),"r"(y));
      }
    } else if (MODE==3){ // mad.hi only
      #pragma unroll
      for(int k2=0;k2<4;k2++){
        #pragma unroll
        for(int i=0;i<32;i++) asm volatile("mad.hi.u32 %0,%1,%2,%0;":"+r"(r[i]):"r"(r[(i+1)&31]),"r"(y));
      }
    } else if (MODE==4){ // mad.wide no carry (64) interleaved with 64 IADD3-ish adds on ALU pipe (128 ops, 64 imad)
      #pragma unroll
      for(int k2=0;k2<8;k2++){
        asm volatile("mad.wide.u32 %0,%1,%2,%0;":"+l"(a0):"r"((u32)a1),"r"(y)); asm volatile("add.cc.u32 %0,%0,%1;":"+r"(r[0]):"r"(x));
        asm volatile("mad.wide.u32 %0,%1,%2,%0;":"+l"(a1):"r"((u32)a2),"r"(y)); asm volatile("addc.cc.u32 %0,%0,%1;":"+r"(r[1]):"r"(x));
        asm volatile("mad.wide.u32 %0,%1,%2,%0;":"+l"(a2):"r"((u32)a3),"r"(y)); asm volatile("addc.cc.u32 %0,%0,%1;":"+r"(r[2]):"r"(x));
        asm volatile("mad.wide.u32 %0,%1,%2,%0;":"+l"(a3):"r"((u32)a4),"r"(y)); asm volatile("addc.cc.u32 %0,%0,%1;":"+r"(r[3]):"r"(x));
        asm volatile("mad.wide.u32 %0,%1,%2,%0;":"+l"(a4):"r"((u32)a5),"r"(y)); asm volatile("addc.cc.u32 %0,%0,%1;":"+r"(r[4]):"r"(x));
        asm volatile("mad.wide.u32 %0,%1,%2,%0;":"+l"(a5):"r"((u32)a6),"r"(y)); asm volatile("addc.cc.u32 %0,%0,%1;":"+r"(r[5]):"r"(x));
        asm volatile("mad.wide.u32 %0,%1,%2,%0;":"+l"(a6):"r"((u32)a7),"r"(y)); asm volatile("addc.cc.u32 %0,%0,%1;":"+r"(r[6]):"r"(x));
        asm volatile("mad.wide.u32 %0,%1,%2,%0;":"+l"(a7):"r"((u32)a0),"r"(y)); asm volatile("addc.u32 %0,%0,%1;":"+r"(r[7]):"r"(x));
      }
    } else if (MODE==7 || MODE==8){ // 8 chains of 16 wide mads per trip = 128; MODE 7 links them through the carry flag
      #pragma unroll
      for(int rep=0;rep<2;rep++){
        if (MODE==7) asm volatile("add.cc.u32 %0,%0,0;":"+r"(r[0]));      // defines CF (=0) for the first link
      #pragma unroll
      for(int c=0;c<4;c++){
        u32* q=&r[8*c];
        if (MODE==7) asm volatile(
            "madc.lo.cc.u32 %0,%8,%9,%0; madc.hi.cc.u32 %1,%8,%9,%1; madc.lo.cc.u32 %2,%8,%9,%2; madc.hi.cc.u32 %3,%8,%9,%3;"
            "madc.lo.cc.u32 %4,%8,%9,%4; madc.hi.cc.u32 %5,%8,%9,%5; madc.lo.cc.u32 %6,%8,%9,%6; madc.hi.cc.u32 %7,%8,%9,%7;"
            "madc.lo.cc.u32 %0,%8,%9,%0; madc.hi.cc.u32 %1,%8,%9,%1; madc.lo.cc.u32 %2,%8,%9,%2; madc.hi.cc.u32 %3,%8,%9,%3;"
            "madc.lo.cc.u32 %4,%8,%9,%4; madc.hi.cc.u32 %5,%8,%9,%5; madc.lo.cc.u32 %6,%8,%9,%6; madc.hi.cc.u32 %7,%8,%9,%7;"
            :"+r"(q[0]),"+r"(q[1]),"+r"(q[2]),"+r"(q[3]),"+r"(q[4]),"+r"(q[5]),"+r"(q[6]),"+r"(q[7]):"r"(r[(8*c+9)&31]),"r"(y));
        else asm volatile(
            "mad.lo.cc.u32 %0,%8,%9,%0; madc.hi.cc.u32 %1,%8,%9,%1; madc.lo.cc.u32 %2,%8,%9,%2; madc.hi.cc.u32 %3,%8,%9,%3;"
            "madc.lo.cc.u32 %4,%8,%9,%4; madc.hi.cc.u32 %5,%8,%9,%5; madc.lo.cc.u32 %6,%8,%9,%6; madc.hi.cc.u32 %7,%8,%9,%7;"
            "madc.lo.cc.u32 %0,%8,%9,%0; madc.hi.cc.u32 %1,%8,%9,%1; madc.lo.cc.u32 %2,%8,%9,%2; madc.hi.cc.u32 %3,%8,%9,%3;"
            "madc.lo.cc.u32 %4,%8,%9,%4; madc.hi.cc.u32 %5,%8,%9,%5; madc.lo.cc.u32 %6,%8,%9,%6; madc.hi.cc.u32 %7,%8,%9,%7;"
            :"+r"(q[0]),"+r"(q[1]),"+r"(q[2]),"+r"(q[3]),"+r"(q[4]),"+r"(q[5]),"+r"(q[6]),"+r"(q[7]):"r"(r[(8*c+9)&31]),"r"(y));
      }}
    } else if (MODE==9 || MODE==10){ // 64 wide mads + 64 (or 128) independent ALU ops (xor/add on other registers)
      u32 e0=r[16],e1=r[17],e2=r[18],e3=r[19],e4=r[20],e5=r[21],e6=r[22],e7=r[23];
      #pragma unroll
      for(int rep=0;rep<2;rep++){
      #pragma unroll
      for(int c=0;c<2;c++){
        u32* q=&r[8*c];
        #pragma unroll
        for(int h2=0;h2<2;h2++){
        asm volatile(
            "mad.lo.cc.u32 %0,%8,%9,%0; madc.hi.cc.u32 %1,%8,%9,%1; madc.lo.cc.u32 %2,%8,%9,%2; madc.hi.cc.u32 %3,%8,%9,%3;"
            "madc.lo.cc.u32 %4,%8,%9,%4; madc.hi.cc.u32 %5,%8,%9,%5; madc.lo.cc.u32 %6,%8,%9,%6; madc.hi.cc.u32 %7,%8,%9,%7;"
            "madc.lo.cc.u32 %0,%8,%9,%0; madc.hi.cc.u32 %1,%8,%9,%1; madc.lo.cc.u32 %2,%8,%9,%2; madc.hi.cc.u32 %3,%8,%9,%3;"
            "madc.lo.cc.u32 %4,%8,%9,%4; madc.hi.cc.u32 %5,%8,%9,%5; madc.lo.cc.u32 %6,%8,%9,%6; madc.hi.u32 %7,%8,%9,%7;"
            :"+r"(q[0]),"+r"(q[1]),"+r"(q[2]),"+r"(q[3]),"+r"(q[4]),"+r"(q[5]),"+r"(q[6]),"+r"(q[7]):"r"(r[24+c]),"r"(y));
        #pragma unroll
        for(int z2=0; z2<(MODE==9?1:2); z2++){
          asm volatile("xor.b32 %0,%0,%1; add.u32 %2,%2,%3; xor.b32 %4,%4,%5; add.u32 %6,%6,%7;":"+r"(e0),"+r"(e1),"+r"(e2),"+r"(e3),"+r"(e4),"+r"(e5),"+r"(e6),"+r"(e7));
          asm volatile("add.u32 %0,%0,%1; xor.b32 %2,%2,%3; add.u32 %4,%4,%5; xor.b32 %6,%6,%7;":"+r"(e1),"+r"(e2),"+r"(e3),"+r"(e4),"+r"(e5),"+r"(e6),"+r"(e7),"+r"(e0));
          asm volatile("xor.b32 %0,%0,%1; add.u32 %2,%2,%3; xor.b32 %4,%4,%5; add.u32 %6,%6,%7;":"+r"(e0),"+r"(e1),"+r"(e2),"+r"(e3),"+r"(e4),"+r"(e5),"+r"(e6),"+r"(e7));
          asm volatile("add.u32 %0,%0,%1; xor.b32 %2,%2,%3; add.u32 %4,%4,%5; xor.b32 %6,%6,%7;":"+r"(e1),"+r"(e2),"+r"(e3),"+r"(e4),"+r"(e5),"+r"(e6),"+r"(e7),"+r"(e0));
        }
        }
      }}
      r[16]=e0;r[17]=e1;r[18]=e2;r[19]=e3;r[20]=e4;r[21]=e5;r[22]=e6;r[23]=e7;
    } else if (MODE==5){ // DFMA: 8 chains x 16
      double d0=__longlong_as_double(a0|0x3ff0000000000000ull),d1=d0+1,d2=d0+2,d3=d0+3,d4=d0+4,d5=d0+5,d6=d0+6,d7=d0+7;
      double fx=1.0000001, fy=1e-9;
      #pragma unroll
      for(int k2=0;k2<16;k2++){ d0=fma(d0,fx,fy);d1=fma(d1,fx,fy);d2=fma(d2,fx,fy);d3=fma(d3,fx,fy);d4=fma(d4,fx,fy);d5=fma(d5,fx,fy);d6=fma(d6,fx,fy);d7=fma(d7,fx,fy);}
      a0^=__double_as_longlong(d0+d1+d2+d3+d4+d5+d6+d7);
    } else if (MODE==6){ // mad.wide no carry + DFMA interleaved (64 + 64)
      double d0=__longlong_as_double(a7|0x3ff0000000000000ull),d1=d0+1,d2=d0+2,d3=d0+3;
      double fx=1.0000001, fy=1e-9;
      #pragma unroll
      for(int k2=0;k2<16;k2++){
        asm volatile("mad.wide.u32 %0,%1,%2,%0;":"+l"(a0):"r"((u32)a1),"r"(y)); d0=fma(d0,fx,fy);
        asm volatile("mad.wide.u32 %0,%1,%2,%0;":"+l"(a1):"r"((u32)a2),"r"(y)); d1=fma(d1,fx,fy);
        asm volatile("mad.wide.u32 %0,%1,%2,%0;":"+l"(a2):"r"((u32)a3),"r"(y)); d2=fma(d2,fx,fy);
        asm volatile("mad.wide.u32 %0,%1,%2,%0;":"+l"(a3):"r"((u32)a4),"r"(y)); d3=fma(d3,fx,fy);
      }
      a7^=__double_as_longlong(d0+d1+d2+d3);
    }
  }
  u64 s=a0^a1^a2^a3^a4^a5^a6^a7; u32 z=0;
  #pragma unroll
  for(int i=0;i<32;i++) z^=r[i];
  if((u32)(s^(s>>32)^z)==0x23456789u) sink[0]=(u32)s;
}
template<int MODE> void run(const char* name, double ops_per_iter, int sms){
  u32* d; cudaMalloc(&d,4); cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int blocks=sms*8, threads=256, iters=2000;
  k<MODE><<<blocks,threads>>>(d,100); cudaDeviceSynchronize();
  float best=1e30f;
  for(int r=0;r<3;r++){ cudaEventRecord(e0); k<MODE><<<blocks,threads>>>(d,iters); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
  double total=(double)blocks*threads*iters*ops_per_iter;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-44s %8.3f ms  %7.2f Tops/s  %6.1f ops/clk/SM (at %d MHz)\n", name, best, total/best/1e9, total/(best*1e-3)/sms/(clk*1e3), clk/1000);
}
int main(){ int sms; cudaDeviceGetAttribute(&sms,cudaDevAttrMultiProcessorCount,0);
  run<0>("mad.wide.u32 (IMAD.WIDE, no carry)",128,sms);
  run<1>("mad.lo.cc+madc.hi.cc pairs (IMAD.WIDE.X)",128,sms);
  run<2>("mad.lo.u32 (IMAD)",128,sms);
  run<3>("mad.hi.u32 (IMAD.HI)",128,sms);
  run<4>("mad.wide + add.cc chain interleaved (64 imad)",64,sms);
  run<5>("DFMA",128,sms);
  run<6>("mad.wide + DFMA interleaved (64 imad+64 dfma)",128,sms);
  run<7>("IMAD.WIDE.X, 4 chains LINKED via carry flag",64,sms);
  run<8>("IMAD.WIDE.X, 4 chains independent asm blocks",64,sms);
  run<9>("IMAD.WIDE.X 64 + 64 independent ALU ops (wide/s)",64,sms);
  run<10>("IMAD.WIDE.X 64 + 128 independent ALU ops (wide/s)",64,sms);
  return 0; }
