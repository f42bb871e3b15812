// measurement aid: dz_value (scalar) against dz_value2 (packed pairs) on random inputs, bit for bit
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include "../3dfeatnet_b200/csrc/dz_source.cuh"
using namespace f3d;
__global__ void k(const float* v, int n, unsigned long long* bad) {
  int i = blockIdx.x*blockDim.x+threadIdx.x; if (i >= n) return;
  const float* p = v + (size_t)i*20;
  for (int relu = 0; relu < 2; ++relu) {
    float pm0 = (i&1) ? __fmaf_rn(p[0],p[2],p[4]) : p[16]; if (relu && (i&1)) pm0 = fmaxf(pm0, 0.f);
    float pm1 = (i&2) ? __fmaf_rn(p[1],p[3],p[5]) : p[17]; if (relu && (i&2)) pm1 = fmaxf(pm1, 0.f);
    float a0 = dz_value(p[0],p[2],p[4],p[6],p[8],p[10],p[12],p[14],pm0,p[18],relu);
    float a1 = dz_value(p[1],p[3],p[5],p[7],p[9],p[11],p[13],p[15],pm1,p[19],relu);
    float b0,b1; dz_value2(p[0],p[1],p[2],p[3],p[4],p[5],p[6],p[7],p[8],p[9],p[10],p[11],p[12],p[13],p[14],p[15],pm0,pm1,p[18],p[19],relu,b0,b1);
    if (__float_as_uint(a0)!=__float_as_uint(b0) || __float_as_uint(a1)!=__float_as_uint(b1)) { if (atomicAdd(bad,1ull)<6) printf("relu %d: %a %a vs %a %a\n", relu,a0,a1,b0,b1); }
  }
}
int main(){ int n=1<<20; float *h=(float*)malloc((size_t)n*20*4); srand(3);
  for (size_t i=0;i<(size_t)n*20;i++) h[i] = (rand()/(float)RAND_MAX - 0.5f) * ((i%20)<2 ? 8.f : 2.f);
  float *d; cudaMalloc(&d,(size_t)n*20*4); cudaMemcpy(d,h,(size_t)n*20*4,cudaMemcpyHostToDevice); unsigned long long *bad; cudaMalloc(&bad,8); cudaMemset(bad,0,8);
  k<<<n/256,256>>>(d,n,bad); unsigned long long hb; cudaMemcpy(&hb,bad,8,cudaMemcpyDeviceToHost);
  printf("dz mismatches %llu of %d (%s)\n",hb,2*n,cudaGetErrorString(cudaGetLastError())); return 0; }
