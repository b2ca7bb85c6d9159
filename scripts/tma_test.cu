#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void* p){ return (uint32_t)__cvta_generic_to_shared(p); }
template<int RANK>
__global__ void k(const __grid_constant__ CUtensorMap tmap, int x, int y, int z, uint8_t* out, int R){
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ uint64_t bar;
  if (threadIdx.x==0){ asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;"::"r"(s32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  if (threadIdx.x==0){
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"::"r"(s32(&bar)),"r"(R*272));
    if (RANK==3) asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"::"r"(s32(sm)),"l"(&tmap),"r"(x),"r"(y),"r"(z),"r"(s32(&bar)):"memory");
    else asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"::"r"(s32(sm)),"l"(&tmap),"r"(x),"r"(y),"r"(s32(&bar)):"memory");
  }
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}"::"r"(s32(&bar)):"memory");
  for (int i=threadIdx.x;i<R*272;i+=blockDim.x) out[i]=sm[i];
}
int main(){
  void* p=nullptr; cudaDriverEntryPointQueryResult qr; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled",&p,cudaEnableDefault,&qr); enc_fn enc=(enc_fn)p; printf("enc %p qr %d\n",p,(int)qr);
  int rows=700, cols=1000, step=1008, pages=2; uint8_t* d; cudaMalloc(&d,(size_t)step*rows*pages); 
  uint8_t* h=(uint8_t*)malloc((size_t)step*rows*pages); for(size_t i=0;i<(size_t)step*rows*pages;i++) h[i]=(uint8_t)(i*7+i/1008); cudaMemcpy(d,h,(size_t)step*rows*pages,cudaMemcpyHostToDevice);
  uint8_t* out; cudaMalloc(&out,272*8); uint8_t ho[2176];
  for (int rank=2; rank<=3; ++rank) for (int R: {8}) for (int x: {0,240,-16,752,992}) {
    CUtensorMap m; cuuint64_t gd[3]={(cuuint64_t)step/2,(cuuint64_t)rows,(cuuint64_t)pages}; cuuint64_t gs[2]={(cuuint64_t)step,(cuuint64_t)step*rows}; cuuint32_t box[3]={136,(cuuint32_t)R,1}; cuuint32_t es[3]={1,1,1};
    CUresult cr=enc(&m,CU_TENSOR_MAP_DATA_TYPE_UINT16,rank,d,gd,gs,box,es,CU_TENSOR_MAP_INTERLEAVE_NONE,CU_TENSOR_MAP_SWIZZLE_NONE,CU_TENSOR_MAP_L2_PROMOTION_L2_128B,CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cudaMemset(out,0xEE,2176);
    if (rank==3) k<3><<<1,64,2304>>>(m,x/2,16,1,out,R); else k<2><<<1,64,2304>>>(m,x/2,16,0,out,R);
    cudaError_t e=cudaDeviceSynchronize(); cudaMemcpy(ho,out,2176,cudaMemcpyDeviceToHost);
    int z=(rank==3)?1:0; int bad=0; for(int r=0;r<R;r++) for(int c=0;c<272;c++){ int xx=x+c; uint8_t want=(xx<0||xx>=step)?0:h[(size_t)z*step*rows+(size_t)(16+r)*step+xx]; if(ho[r*272+c]!=want) bad++; }
    printf("rank %d R %d x %d: enc %d run %s bad %d\n",rank,R,x,(int)cr,cudaGetErrorString(e),bad);
    if (e!=cudaSuccess) return 1;
  }
  return 0; }
