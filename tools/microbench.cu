// tools/microbench.cu — B200 micro-measurements that decide the sweep design:
// match.any cost, shared-memory atomic vs plain-store throughput, L2-resident
// random gather rate.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d\n",cudaGetErrorString(e),__LINE__);return 1;}}while(0)

__device__ __forceinline__ uint32_t mix(uint32_t x){x^=x>>16;x*=0x7feb352dU;x^=x>>15;x*=0x846ca68bU;x^=x>>16;return x;}

template<int MODE> __global__ void k_match(int iters, unsigned* out){
  unsigned lane=threadIdx.x&31, acc=0; uint32_t s=mix(blockIdx.x*blockDim.x+threadIdx.x+1);
  for(int i=0;i<iters;i++){
    s=s*1664525u+1013904223u;
    uint32_t key = MODE==0 ? s /*all distinct (whp)*/ : MODE==1 ? (s>>16&0) + (uint32_t)i /*all equal*/ : ((s>>13)&3)+(uint32_t)i*4 /*4 groups*/;
    acc += __match_any_sync(0xffffffffu,key);
  }
  if(acc==0x12345u) out[0]=acc+lane;
}
template<int MODE> __global__ void k_smem(int iters, unsigned* out){
  extern __shared__ unsigned sm[]; const int N=8192; // 32 KB
  for(int i=threadIdx.x;i<N;i+=blockDim.x) sm[i]=0; __syncthreads();
  uint32_t s=mix(blockIdx.x*blockDim.x+threadIdx.x+1); unsigned acc=0;
  for(int i=0;i<iters;i++){
    s=s*1664525u+1013904223u; unsigned idx=(s>>8)&(N-1);
    if(MODE==0) atomicAdd(&sm[idx],1u);               // ATOMS/RED no return
    else if(MODE==1) acc+=atomicAdd(&sm[idx],1u);     // with return
    else if(MODE==2) acc+=atomicCAS(&sm[idx],0u,s);   // CAS
    else if(MODE==3) acc+=sm[idx];                    // LDS random
    else if(MODE==4) sm[idx]=s;                       // STS random
    else if(MODE==5){ unsigned v=sm[idx]; sm[idx]=v+1; } // LDS+STS
  }
  __syncthreads(); if(acc==0x12345u||sm[threadIdx.x]==0xdeadbeef) out[0]=acc;
}
__global__ void k_smem_f64(int iters, double* out){
  extern __shared__ double smd[]; const int N=4096;
  for(int i=threadIdx.x;i<N;i+=blockDim.x) smd[i]=0; __syncthreads();
  uint32_t s=mix(blockIdx.x*blockDim.x+threadIdx.x+1);
  for(int i=0;i<iters;i++){ s=s*1664525u+1013904223u; atomicAdd(&smd[(s>>8)&(N-1)],1.0); }
  __syncthreads(); if(smd[threadIdx.x]==-1.0) out[0]=1;
}
// random 4B gather from a table of n ints (power of two), U independent loads in flight
template<int U> __global__ void k_gather(const int* __restrict__ tab, uint32_t mask, int iters, unsigned* out){
  uint32_t s=mix(blockIdx.x*blockDim.x+threadIdx.x+1); unsigned acc=0;
  for(int i=0;i<iters;i++){
    int v[U];
    #pragma unroll
    for(int u=0;u<U;u++){ s=s*1664525u+1013904223u; v[u]=tab[(s>>4)&mask]; }
    #pragma unroll
    for(int u=0;u<U;u++) acc+=v[u];
  }
  if(acc==0x12345u) out[0]=acc;
}
// skewed gather: index = product of two randoms >> shift (concentrates on low ids)
template<int U> __global__ void k_gather_skew(const int* __restrict__ tab, uint32_t mask, int iters, unsigned* out){
  uint32_t s=mix(blockIdx.x*blockDim.x+threadIdx.x+1); unsigned acc=0;
  for(int i=0;i<iters;i++){
    int v[U];
    #pragma unroll
    for(int u=0;u<U;u++){ s=s*1664525u+1013904223u; uint32_t a=(s>>4)&mask; s=s*1664525u+1013904223u; uint32_t b=(s>>4)&mask; s=s*1664525u+1013904223u; uint32_t c=(s>>4)&mask; v[u]=tab[a&b&c]; }
    #pragma unroll
    for(int u=0;u<U;u++) acc+=v[u];
  }
  if(acc==0x12345u) out[0]=acc;
}
__global__ void k_red_f64(double* tab, uint32_t mask, int iters){
  uint32_t s=mix(blockIdx.x*blockDim.x+threadIdx.x+1);
  for(int i=0;i<iters;i++){ s=s*1664525u+1013904223u; atomicAdd(&tab[(s>>4)&mask],1.0); }
}
__global__ void k_stream(const int4* __restrict__ a, size_t n, unsigned* out){
  unsigned acc=0; for(size_t i=blockIdx.x*(size_t)blockDim.x+threadIdx.x;i<n;i+=(size_t)gridDim.x*blockDim.x){int4 v=a[i];acc+=v.x+v.y+v.z+v.w;}
  if(acc==0x12345u) out[0]=acc;
}
template<typename F> float timeit(F f,int rep=5){cudaEvent_t a,b;cudaEventCreate(&a);cudaEventCreate(&b);f();cudaDeviceSynchronize();float best=1e30f;for(int r=0;r<rep;r++){cudaEventRecord(a);f();cudaEventRecord(b);cudaEventSynchronize(b);float ms;cudaEventElapsedTime(&ms,a,b);if(ms<best)best=ms;}return best;}
int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0)); int sm=p.multiProcessorCount; int clk; cudaDeviceGetAttribute(&clk,cudaDevAttrClockRate,0);
  printf("device %s SMs %d clock %.0f MHz L2 %d MB\n",p.name,sm,clk/1e3,p.l2CacheSize>>20);
  unsigned* out; CK(cudaMalloc(&out,64)); const int B=256, G=sm*8, IT=4096; double lanes=(double)G*B*IT;
  #define REP(name,call) { float ms=timeit([&]{call;}); CK(cudaGetLastError()); printf("%-34s %8.3f ms  %8.2f Glane-ops/s  %6.3f lane-ops/clk/SM\n",name,ms,lanes/ms/1e6,lanes/(ms*1e-3)/(sm*(double)clk*1e3)); }
  REP("match_any distinct", (k_match<0><<<G,B>>>(IT,out)));
  REP("match_any all-equal", (k_match<1><<<G,B>>>(IT,out)));
  REP("match_any 4 groups", (k_match<2><<<G,B>>>(IT,out)));
  cudaFuncSetAttribute(k_smem<0>,cudaFuncAttributeMaxDynamicSharedMemorySize,32768);
  REP("smem atomicAdd u32 noret", (k_smem<0><<<G,B,32768>>>(IT,out)));
  REP("smem atomicAdd u32 ret", (k_smem<1><<<G,B,32768>>>(IT,out)));
  REP("smem atomicCAS u32", (k_smem<2><<<G,B,32768>>>(IT,out)));
  REP("smem LDS random", (k_smem<3><<<G,B,32768>>>(IT,out)));
  REP("smem STS random", (k_smem<4><<<G,B,32768>>>(IT,out)));
  REP("smem LDS+STS rmw", (k_smem<5><<<G,B,32768>>>(IT,out)));
  REP("smem atomicAdd f64", (k_smem_f64<<<G,B,32768>>>(IT,(double*)out)));
  for(int mb : {4, 64, 512}){
    size_t n=(size_t)mb<<18; int* tab; CK(cudaMalloc(&tab,n*4)); CK(cudaMemset(tab,1,n*4));
    char nm[64]; const int IT2=512; double l2=(double)G*B*IT2*8; 
    { float ms=timeit([&]{k_gather<8><<<G,B>>>(tab,(uint32_t)n-1,IT2,out);}); snprintf(nm,64,"gather uniform %3d MB",mb); printf("%-34s %8.3f ms  %8.2f Ggather/s  (%.1f GB/s of 32B sectors)\n",nm,ms,l2/ms/1e6,l2*32/ms/1e6);}
    { float ms=timeit([&]{k_gather_skew<8><<<G,B>>>(tab,(uint32_t)n-1,IT2,out);}); snprintf(nm,64,"gather skewed  %3d MB",mb); printf("%-34s %8.3f ms  %8.2f Ggather/s\n",nm,ms,l2/ms/1e6);}
    { float ms=timeit([&]{k_red_f64<<<G,B>>>((double*)tab,(uint32_t)(n/2)-1,IT2);}); snprintf(nm,64,"RED.f64 random %3d MB",mb); printf("%-34s %8.3f ms  %8.2f Gred/s\n",nm,ms,(double)G*B*IT2/ms/1e6);}
    cudaFree(tab);
  }
  { size_t n=(size_t)1<<30; int4* a; CK(cudaMalloc(&a,n)); CK(cudaMemset(a,0,n)); float ms=timeit([&]{k_stream<<<sm*16,512>>>(a,n/16,out);}); printf("%-34s %8.3f ms  %8.1f GB/s\n","stream read 1 GiB",ms,n/ms/1e6); cudaFree(a);}
  return 0;
}
