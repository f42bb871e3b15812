#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
__device__ __forceinline__ uint64_t pk2(float a, float b){uint64_t r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ void upk2(uint64_t v, float&a, float&b){asm("mov.b64 {%0,%1}, %2;":"=f"(a),"=f"(b):"l"(v));}
__global__ void k(const float* a, const float* b, const float* c, int n, unsigned long long* bad) {
  int i = blockIdx.x*blockDim.x+threadIdx.x; if (2*i+1 >= n) return;
  float a0=a[2*i],a1=a[2*i+1],b0=b[2*i],b1=b[2*i+1],c0=c[2*i],c1=c[2*i+1];
  uint64_t r; float x0,x1;
  asm("add.rn.f32x2 %0, %1, %2;":"=l"(r):"l"(pk2(a0,a1)),"l"(pk2(b0,b1))); upk2(r,x0,x1);
  if (__float_as_uint(x0)!=__float_as_uint(__fadd_rn(a0,b0)) || __float_as_uint(x1)!=__float_as_uint(__fadd_rn(a1,b1))) atomicAdd(&bad[0],1ull);
  asm("sub.rn.f32x2 %0, %1, %2;":"=l"(r):"l"(pk2(a0,a1)),"l"(pk2(b0,b1))); upk2(r,x0,x1);
  if (__float_as_uint(x0)!=__float_as_uint(__fsub_rn(a0,b0)) || __float_as_uint(x1)!=__float_as_uint(__fsub_rn(a1,b1))) atomicAdd(&bad[1],1ull);
  asm("mul.rn.f32x2 %0, %1, %2;":"=l"(r):"l"(pk2(a0,a1)),"l"(pk2(b0,b1))); upk2(r,x0,x1);
  if (__float_as_uint(x0)!=__float_as_uint(__fmul_rn(a0,b0)) || __float_as_uint(x1)!=__float_as_uint(__fmul_rn(a1,b1))) { if (atomicAdd(&bad[2],1ull) < 4) printf("mul %a * %a = %a vs %a | %a * %a = %a vs %a\n", a0,b0,x0,__fmul_rn(a0,b0),a1,b1,x1,__fmul_rn(a1,b1)); }
  asm("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(r):"l"(pk2(a0,a1)),"l"(pk2(b0,b1)),"l"(pk2(c0,c1))); upk2(r,x0,x1);
  if (__float_as_uint(x0)!=__float_as_uint(__fmaf_rn(a0,b0,c0)) || __float_as_uint(x1)!=__float_as_uint(__fmaf_rn(a1,b1,c1))) atomicAdd(&bad[3],1ull);
}
int main(){ int n=1<<24; float *h=(float*)malloc(3*n*4); srand(1);
  for (int i=0;i<3*n;i++){ uint32_t u=((uint32_t)rand()<<16)^rand(); int m=i%7; if (m==0) u&=0x807fffff; /* subnormal */ if (m==1) u=(u&0x80000000); if (m==2) u=(u&0x807fffff)|0x00800000u*(1+(u>>27)%4); memcpy(&h[i],&u,4);} 
  float *d; cudaMalloc(&d,3*n*4); cudaMemcpy(d,h,3*n*4,cudaMemcpyHostToDevice); unsigned long long *bad; cudaMalloc(&bad,32); cudaMemset(bad,0,32);
  k<<<n/2/256,256>>>(d,d+n,d+2*n,n,bad); unsigned long long hb[4]; cudaMemcpy(hb,bad,32,cudaMemcpyDeviceToHost);
  printf("mismatches add %llu sub %llu mul %llu fma %llu of %d pairs (%s)\n",hb[0],hb[1],hb[2],hb[3],n/2,cudaGetErrorString(cudaGetLastError())); return 0; }
