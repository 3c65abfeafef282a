// times the fixed per-process CUDA costs of the drop-in executable: init, context, big allocations, teardown
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>
#include <time.h>
static double now(){struct timespec t; clock_gettime(CLOCK_MONOTONIC,&t); return t.tv_sec+t.tv_nsec*1e-9;}
int main(int argc,char**argv){
  double t0=now(); int hold = argc>1 ? atoi(argv[1]) : 0; size_t gb = argc>2 ? atol(argv[2]) : 30;
  cuInit(0); double t1=now();
  int n=0; cuDeviceGetCount(&n); CUdevice d; cuDeviceGet(&d,0); double t2=now();
  cudaSetDevice(0); cudaFree(0); double t3=now();
  void *p[64]; for(size_t i=0;i<gb;i++) cudaMalloc(&p[i],(size_t)1<<30); double t4=now();
  void *h; cudaHostAlloc(&h,(size_t)128<<20,cudaHostAllocDefault); double t5=now();
  cudaMemset(p[0],0,(size_t)1<<30); cudaDeviceSynchronize(); double t6=now();
  printf("devices %d | cuInit %.3f  devget %.3f  ctx %.3f  malloc %zuGB %.3f  hostalloc128MB %.3f  memset %.3f\n",n,t1-t0,t2-t1,t3-t2,gb,t4-t3,t5-t4,t6-t5);
  fflush(stdout);
  if(hold){ sleep(hold); return 0; }
  double t7=now(); for(size_t i=0;i<gb;i++) cudaFree(p[i]); cudaFreeHost(h); double t8=now();
  cudaDeviceReset(); double t9=now();
  printf("free %.3f  reset %.3f\n",t8-t7,t9-t8);
  return 0;}
