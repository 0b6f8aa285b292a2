#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double &c0,double &c1,double a,double b){
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n":"+d"(c0),"+d"(c1):"d"(a),"d"(b));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]){
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
   :"+d"(c[0]),"+d"(c[1]),"+d"(c[2]),"+d"(c[3])
   :"d"(a[0]),"d"(a[1]),"d"(a[2]),"d"(a[3]),"d"(a[4]),"d"(a[5]),"d"(a[6]),"d"(a[7]),"d"(b[0]),"d"(b[1]),"d"(b[2]),"d"(b[3]));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]){
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
   :"+d"(c[0]),"+d"(c[1]),"+d"(c[2]),"+d"(c[3])
   :"d"(a[0]),"d"(a[1]),"d"(a[2]),"d"(a[3]),"d"(b[0]),"d"(b[1]));
}
__device__ __forceinline__ void dmma1684(double (&c)[4], const double (&a)[2], const double (&b)[1]){
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
   :"+d"(c[0]),"+d"(c[1]),"+d"(c[2]),"+d"(c[3])
   :"d"(a[0]),"d"(a[1]),"d"(b[0]));
}
template<int NACC> __global__ void k884(double* out, int iters, double a, double b){
  double c0[NACC], c1[NACC];
  for(int i=0;i<NACC;i++){c0[i]=threadIdx.x+i;c1[i]=i;}
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<NACC;i++) dmma884(c0[i],c1[i],a,b);
  }
  double s=0; for(int i=0;i<NACC;i++) s+=c0[i]+c1[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int NACC> __global__ void k16816(double* out, int iters, double av, double bv){
  double c[NACC][4]; double a[8], b[4];
  for(int i=0;i<8;i++) a[i]=av+i; for(int i=0;i<4;i++) b[i]=bv+i;
  for(int i=0;i<NACC;i++) for(int j=0;j<4;j++) c[i][j]=threadIdx.x+i+j;
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<NACC;i++) dmma16816(c[i],a,b);
  }
  double s=0; for(int i=0;i<NACC;i++) for(int j=0;j<4;j++) s+=c[i][j];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int NACC> __global__ void k1688(double* out, int iters, double av, double bv){
  double c[NACC][4]; double a[4], b[2];
  for(int i=0;i<4;i++) a[i]=av+i; for(int i=0;i<2;i++) b[i]=bv+i;
  for(int i=0;i<NACC;i++) for(int j=0;j<4;j++) c[i][j]=threadIdx.x+i+j;
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<NACC;i++) dmma1688(c[i],a,b);
  }
  double s=0; for(int i=0;i<NACC;i++) for(int j=0;j<4;j++) s+=c[i][j];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int NACC> __global__ void k1684(double* out, int iters, double av, double bv){
  double c[NACC][4]; double a[2], b[1];
  for(int i=0;i<2;i++) a[i]=av+i; b[0]=bv;
  for(int i=0;i<NACC;i++) for(int j=0;j<4;j++) c[i][j]=threadIdx.x+i+j;
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<NACC;i++) dmma1684(c[i],a,b);
  }
  double s=0; for(int i=0;i<NACC;i++) for(int j=0;j<4;j++) s+=c[i][j];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int NACC> __global__ void kfma(double* out, int iters, double a, double b){
  double c[NACC];
  for(int i=0;i<NACC;i++) c[i]=threadIdx.x+i;
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<NACC;i++) c[i]=fma(c[i],a,b);
  }
  double s=0; for(int i=0;i<NACC;i++) s+=c[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<typename F> float timeit(F f){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best=1e30;
  for(int r=0;r<5;r++){ cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms;}
  return best;
}
int main(){
  double* out; cudaMalloc(&out, 148*8*1024*sizeof(double));
  int iters=20000;
  for(int wpb : {2,4,8,16}){
    int threads=wpb*32; int blocks=148*(wpb<=4?1:2);
    double nw=(double)blocks*wpb;
    float ms;
    ms=timeit([&]{k884<16><<<blocks,threads>>>(out,iters,1.0000001,0.999999);});
    printf("m8n8k4   warps/blk %2d blocks %d: %.2f TFLOP/s\n",wpb,blocks, nw*iters*16*512.0/ms/1e9);
    ms=timeit([&]{k1684<8><<<blocks,threads>>>(out,iters,1.0000001,0.999999);});
    printf("m16n8k4  warps/blk %2d blocks %d: %.2f TFLOP/s\n",wpb,blocks, nw*iters*8*1024.0/ms/1e9);
    ms=timeit([&]{k1688<8><<<blocks,threads>>>(out,iters,1.0000001,0.999999);});
    printf("m16n8k8  warps/blk %2d blocks %d: %.2f TFLOP/s\n",wpb,blocks, nw*iters*8*2048.0/ms/1e9);
    ms=timeit([&]{k16816<8><<<blocks,threads>>>(out,iters/2,1.0000001,0.999999);});
    printf("m16n8k16 warps/blk %2d blocks %d: %.2f TFLOP/s\n",wpb,blocks, nw*(iters/2)*8*4096.0/ms/1e9);
    ms=timeit([&]{kfma<16><<<blocks,threads>>>(out,iters,1.0000001,0.999999);});
    printf("DFMA     warps/blk %2d blocks %d: %.2f TFLOP/s\n",wpb,blocks, nw*iters*16*64.0/ms/1e9);
  }
  cudaError_t e=cudaGetLastError(); printf("err=%s\n",cudaGetErrorString(e));
  return 0;
}
