// Calibration: what plain streaming kernels reach on this B200 (write-only 16B/32B, read-only, copy).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void w16(int4* p, size_t n){ size_t i=blockIdx.x*(size_t)blockDim.x+threadIdx.x; size_t st=(size_t)gridDim.x*blockDim.x; int4 v=make_int4(1,2,3,4); for(;i<n;i+=st) p[i]=v; }
__global__ void w32(longlong4* p, size_t n){ size_t i=blockIdx.x*(size_t)blockDim.x+threadIdx.x; size_t st=(size_t)gridDim.x*blockDim.x; for(;i<n;i+=st){ long long a=i; asm volatile("st.global.v4.s64 [%0], {%1,%1,%1,%1};"::"l"(p+i),"l"(a):"memory"); } }
__global__ void r16(const int4* p, size_t n, int* out){ size_t i=blockIdx.x*(size_t)blockDim.x+threadIdx.x; size_t st=(size_t)gridDim.x*blockDim.x; int acc=0; for(;i<n;i+=st){ int4 v=__ldg(p+i); acc+=v.x^v.y^v.z^v.w; } if(acc==0x12345678) *out=acc; }
__global__ void r32(const longlong4* p, size_t n, int* out){ size_t i=blockIdx.x*(size_t)blockDim.x+threadIdx.x; size_t st=(size_t)gridDim.x*blockDim.x; long long acc=0; for(;i<n;i+=st){ long long a,b,c,d; asm volatile("ld.global.nc.v4.s64 {%0,%1,%2,%3}, [%4];":"=l"(a),"=l"(b),"=l"(c),"=l"(d):"l"(p+i)); acc+=a^b^c^d; } if(acc==0x12345678) *out=(int)acc; }
__global__ void cp16(const int4* s, int4* d, size_t n){ size_t i=blockIdx.x*(size_t)blockDim.x+threadIdx.x; size_t st=(size_t)gridDim.x*blockDim.x; for(;i<n;i+=st) d[i]=__ldg(s+i); }
template<class F> float timeit(F f){ cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b); f(); f(); cudaEventRecord(a); for(int i=0;i<5;i++) f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms,a,b); return ms/5; }
int main(){ size_t bytes=(size_t)16<<30; void *p,*q; cudaMalloc(&p,bytes); cudaMalloc(&q,bytes); int* o; cudaMalloc(&o,4); cudaMemset(p,1,bytes);
  for(int mult: {2,4,8,16}){ int g=148*mult;
   float t;
   t=timeit([&]{w16<<<g,512>>>((int4*)p,bytes/16);}); printf("grid %5d write16 %.0f GB/s\n",g,bytes/t/1e6);
   t=timeit([&]{w32<<<g,512>>>((longlong4*)p,bytes/32);}); printf("grid %5d write32 %.0f GB/s\n",g,bytes/t/1e6);
   t=timeit([&]{r16<<<g,512>>>((int4*)p,bytes/16,o);}); printf("grid %5d read16  %.0f GB/s\n",g,bytes/t/1e6);
   t=timeit([&]{r32<<<g,512>>>((longlong4*)p,bytes/32,o);}); printf("grid %5d read32  %.0f GB/s\n",g,bytes/t/1e6);
   t=timeit([&]{cp16<<<g,512>>>((int4*)p,(int4*)q,bytes/16);}); printf("grid %5d copy16  %.0f GB/s (r+w)\n",g,2*bytes/t/1e6);
  }
  // mixed like K1: 1 byte read per 16 written; like K2: 17 read per 1 written -> approximated by the above
  return 0; }
