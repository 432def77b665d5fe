// Probe (not product code): measures the FP64 DMMA.8x8x4 and DFMA issue rates on B200 and
// verifies the TMA SWIZZLE_64B shared-memory layout formula used by the FP64 GEMM kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_peak dmma_peak.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__device__ __forceinline__ void dmma(double& c0,double& c1,double a,double b){
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};":"+d"(c0),"+d"(c1):"d"(a),"d"(b));
}

template<int NA,int NB>
__global__ void __launch_bounds__(1024) dmma_loop(double* out,int iters){
  double a[NA],b[NB],c[NA*NB*2];
  for(int i=0;i<NA;i++)a[i]=threadIdx.x*1e-9+i;
  for(int i=0;i<NB;i++)b[i]=threadIdx.x*1e-9-i;
  for(int i=0;i<NA*NB*2;i++)c[i]=0;
  for(int it=0;it<iters;it++){
#pragma unroll
    for(int i=0;i<NA;i++)
#pragma unroll
      for(int j=0;j<NB;j++) dmma(c[(i*NB+j)*2],c[(i*NB+j)*2+1],a[i],b[j]);
  }
  double s=0; for(int i=0;i<NA*NB*2;i++)s+=c[i];
  if(s==123.456)out[0]=s;
}

template<int N>
__global__ void __launch_bounds__(1024) dfma_loop(double* out,int iters){
  double c[N]; double a=threadIdx.x*1e-9, b=1.0000001;
  for(int i=0;i<N;i++)c[i]=i;
  for(int it=0;it<iters;it++){
#pragma unroll
    for(int i=0;i<N;i++) c[i]=fma(c[i],b,a);
  }
  double s=0; for(int i=0;i<N;i++)s+=c[i];
  if(s==123.456)out[0]=s;
}

template<typename F> float timeit(F f,int reps=3){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); CK(cudaDeviceSynchronize());
  float best=1e30f;
  for(int r=0;r<reps;r++){ cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
  return best;
}

// ---- TMA swizzle check ----
__global__ void tma_dump(const __grid_constant__ CUtensorMap map, double* out, int c0,int c1,int c2, int nbytes, int rank){
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  uint32_t bar_a=(uint32_t)__cvta_generic_to_shared(&bar);
  uint32_t dst=(uint32_t)__cvta_generic_to_shared(smem);
  if(threadIdx.x==0){
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;"::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if(threadIdx.x==0){
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"::"r"(bar_a),"r"(nbytes));
    if(rank==3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3,%4,%5}], [%2];"
        ::"r"(dst),"l"(&map),"r"(bar_a),"r"(c0),"r"(c1),"r"(c2):"memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3,%4}], [%2];"
        ::"r"(dst),"l"(&map),"r"(bar_a),"r"(c0),"r"(c1):"memory");
  }
  uint32_t ok=0;
  while(!ok){
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p; }":"=r"(ok):"r"(bar_a):"memory");
  }
  for(int i=threadIdx.x;i<nbytes/8;i+=blockDim.x) out[i]=((double*)smem)[i];
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(){
  int dev=0; cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,dev));
  printf("device %s SMs %d clock %d kHz\n",p.name,p.multiProcessorCount,p.clockRate);
  double* out; CK(cudaMalloc(&out,1<<20));
  int nsm=p.multiProcessorCount;
  // DMMA: vary warps per SM and accumulator tiles
  {
    int iters=20000;
    auto run=[&](const char* name,auto kern,int nacc,int warps,int blocks_per_sm){
      float ms=timeit([&]{kern<<<nsm*blocks_per_sm,warps*32>>>(out,iters);});
      double flops=(double)nsm*blocks_per_sm*warps*iters*nacc*512.0; // 8*8*4*2 flops per DMMA
      printf("DMMA %-10s warps/blk %2d blk/SM %d : %.3f ms  %.2f TFLOP/s  (%.2f cyc/DMMA/SMSP @1965MHz)\n",name,warps,blocks_per_sm,ms,flops/ms*1e-9,
        ms*1e-3*1.965e9/((double)blocks_per_sm*warps/4*iters*nacc));
    };
    run("8x4",dmma_loop<8,4>,32,4,1);
    run("8x4",dmma_loop<8,4>,32,8,1);
    run("8x4",dmma_loop<8,4>,32,16,1);
    run("4x4",dmma_loop<4,4>,16,4,1);
    run("4x4",dmma_loop<4,4>,16,8,1);
    run("4x4",dmma_loop<4,4>,16,16,1);
    run("2x2",dmma_loop<2,2>,4,4,1);
    run("2x2",dmma_loop<2,2>,4,8,1);
    run("2x2",dmma_loop<2,2>,4,16,1);
    run("2x2",dmma_loop<2,2>,4,32,1);
    run("1x1",dmma_loop<1,1>,1,4,1);
    run("1x1",dmma_loop<1,1>,1,32,1);
    run("4x7",dmma_loop<4,7>,28,8,1);
  }
  {
    int iters=20000;
    auto run=[&](const char* name,auto kern,int n,int warps){
      float ms=timeit([&]{kern<<<nsm,warps*32>>>(out,iters);});
      double flops=(double)nsm*warps*32*iters*n*2.0;
      printf("DFMA %-6s warps/blk %2d : %.3f ms  %.2f TFLOP/s\n",name,warps,ms,flops/ms*1e-9);
    };
    run("x16",dfma_loop<16>,16,4);
    run("x16",dfma_loop<16>,16,8);
    run("x16",dfma_loop<16>,16,16);
    run("x16",dfma_loop<16>,16,32);
  }
  // ---- TMA swizzle verification ----
  {
    EncodeFn enc=nullptr; cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled",(void**)&enc,cudaEnableDefault,&qr));
    const int LD=256, COLS=64;           // column-major matrix LD x COLS of doubles, value = row + 1000*col
    std::vector<double> h((size_t)LD*COLS);
    for(int c=0;c<COLS;c++)for(int r=0;r<LD;r++)h[(size_t)c*LD+r]=r+1000.0*c;
    double* g; CK(cudaMalloc(&g,h.size()*8)); CK(cudaMemcpy(g,h.data(),h.size()*8,cudaMemcpyHostToDevice));
    double* dump; CK(cudaMalloc(&dump,65536)); std::vector<double> hd(8192);
    CK(cudaFuncSetAttribute(tma_dump,cudaFuncAttributeMaxDynamicSharedMemorySize,65536));
    // (1) MN-major 3D map: d0=8 rows (inner), d1=K columns (stride LD*8), d2=row-blocks (stride 64B). box {8,16,16}
    {
      CUtensorMap m; cuuint64_t dims[3]={8,(cuuint64_t)COLS,(cuuint64_t)LD/8}; cuuint64_t strides[2]={(cuuint64_t)LD*8,64};
      cuuint32_t box[3]={8,16,16}; cuuint32_t es[3]={1,1,1};
      CUresult r=enc(&m,CU_TENSOR_MAP_DATA_TYPE_FLOAT64,3,g,dims,strides,box,es,CU_TENSOR_MAP_INTERLEAVE_NONE,CU_TENSOR_MAP_SWIZZLE_64B,CU_TENSOR_MAP_L2_PROMOTION_L2_128B,CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      printf("encode MN-major 3D: %d\n",(int)r);
      int k0=16, mb0=2; int nbytes=8*16*16*8;
      tma_dump<<<1,128,65536>>>(m,dump,0,k0,mb0,nbytes,3); CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(hd.data(),dump,nbytes,cudaMemcpyDeviceToHost));
      int bad=0;
      for(int mb=0;mb<16;mb++)for(int k=0;k<16;k++)for(int mi=0;mi<8;mi++){
        int off=(mb*16+k)*64 + ((((mi>>1)^((k>>1)&3))<<4)) + (mi&1)*8;   // bytes
        double want=(mb0+mb)*8+mi + 1000.0*(k0+k);
        if(hd[off/8]!=want){ if(bad<5)printf("  mismatch mb%d k%d mi%d got %.0f want %.0f\n",mb,k,mi,hd[off/8],want); bad++; }
      }
      printf("MN-major SW64 formula mismatches: %d / %d\n",bad,16*16*8);
    }
    // (2) K-major 3D map: operand stored with k contiguous: matrix [k rows (contig)][n cols], d0=8 k, d1=N cols (stride LD*8), d2 = k-blocks (stride 64B); box {8,128->64,2}
    {
      CUtensorMap m; cuuint64_t dims[3]={8,(cuuint64_t)COLS,(cuuint64_t)LD/8}; cuuint64_t strides[2]={(cuuint64_t)LD*8,64};
      cuuint32_t box[3]={8,64,2}; cuuint32_t es[3]={1,1,1};
      CUresult r=enc(&m,CU_TENSOR_MAP_DATA_TYPE_FLOAT64,3,g,dims,strides,box,es,CU_TENSOR_MAP_INTERLEAVE_NONE,CU_TENSOR_MAP_SWIZZLE_64B,CU_TENSOR_MAP_L2_PROMOTION_L2_128B,CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      printf("encode K-major 3D: %d\n",(int)r);
      int kb0=4; int nbytes=8*64*2*8;
      tma_dump<<<1,128,65536>>>(m,dump,0,0,kb0,nbytes,3); CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(hd.data(),dump,nbytes,cudaMemcpyDeviceToHost));
      int bad=0;
      for(int kb=0;kb<2;kb++)for(int n=0;n<64;n++)for(int ki=0;ki<8;ki++){
        int off=(kb*64+n)*64 + ((((ki>>1)^((n>>1)&3))<<4)) + (ki&1)*8;
        double want=(kb0+kb)*8+ki + 1000.0*n;
        if(hd[off/8]!=want){ if(bad<5)printf("  mismatch kb%d n%d ki%d got %.0f want %.0f\n",kb,n,ki,hd[off/8],want); bad++; }
      }
      printf("K-major SW64 formula mismatches: %d / %d\n",bad,2*64*8);
    }
    // (3) OOB behaviour: box straddling the end of d1 (columns) -> zero fill expected
    {
      CUtensorMap m; cuuint64_t dims[3]={8,(cuuint64_t)COLS,(cuuint64_t)LD/8}; cuuint64_t strides[2]={(cuuint64_t)LD*8,64};
      cuuint32_t box[3]={8,16,16}; cuuint32_t es[3]={1,1,1};
      enc(&m,CU_TENSOR_MAP_DATA_TYPE_FLOAT64,3,g,dims,strides,box,es,CU_TENSOR_MAP_INTERLEAVE_NONE,CU_TENSOR_MAP_SWIZZLE_64B,CU_TENSOR_MAP_L2_PROMOTION_L2_128B,CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      int nbytes=8*16*16*8;
      tma_dump<<<1,128,65536>>>(m,dump,0,COLS-8,24,nbytes,3); CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(hd.data(),dump,nbytes,cudaMemcpyDeviceToHost));
      int nz=0,zero_expected=0,badz=0;
      for(int mb=0;mb<16;mb++)for(int k=0;k<16;k++)for(int mi=0;mi<8;mi++){
        int off=(mb*16+k)*64 + ((((mi>>1)^((k>>1)&3))<<4)) + (mi&1)*8;
        bool oob=(k>=8)||(24+mb>=LD/8);
        if(oob){zero_expected++; if(hd[off/8]!=0.0)badz++;} else nz++;
      }
      printf("OOB zero-fill: %d oob elements, %d not zero; in-bounds %d\n",zero_expected,badz,nz);
    }
  }
  printf("done\n");
  return 0;
}
