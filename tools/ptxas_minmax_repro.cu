#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
__global__ void k(const uint8_t *__restrict__ pad, int w, int h, int pitch, uint8_t *__restrict__ score)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const uint8_t *c = pad + (size_t)(y + 19) * pitch + x + 19;
    const int v = c[0];
    const int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
    const int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
    int d[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) d[k] = v - (int)c[dy[k] * pitch + dx[k]];
    int best = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        int mn = d[k], mx = d[k];
#pragma unroll
        for (int j = 1; j < 9; ++j) { int t = d[(k + j) & 15]; mn = min(mn, t); mx = max(mx, t); }
        best = max(best, max(mn, -mx));
    }
    score[(size_t)y * w + x] = (uint8_t)best;
}

__global__ void kV0(const uint8_t *__restrict__ pad, int w, int h, int pitch, uint8_t *__restrict__ score)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const uint8_t *c = pad + (size_t)(y + 19) * pitch + x + 19;
    const int v = c[0];
    const int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
    const int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
    int d[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) d[k] = v - (int)c[dy[k] * pitch + dx[k]];
    int bb = -255, bd = 255;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        int mn = d[k], mx = d[k];
#pragma unroll
        for (int j = 1; j < 9; ++j) { int t = d[(k + j) & 15]; mn = min(mn, t); mx = max(mx, t); }
        bb = max(bb, mn); bd = min(bd, mx);
    }
    int best = max(0, max(bb, -bd));
    score[(size_t)y * w + x] = (uint8_t)best;
}

__global__ void kV1(const uint8_t *__restrict__ pad, int w, int h, int pitch, uint8_t *__restrict__ score)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const uint8_t *c = pad + (size_t)(y + 19) * pitch + x + 19;
    const int v = c[0];
    const int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
    const int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
    int d[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) d[k] = v - (int)c[dy[k] * pitch + dx[k]];
    int best = 0;
    for (int k = 0; k < 16; ++k) {
        int mn = 255, mx = -255;
        for (int j = 0; j < 9; ++j) { int t = d[(k + j) & 15]; mn = t < mn ? t : mn; mx = t > mx ? t : mx; }
        int s = mn > -mx ? mn : -mx;
        best = s > best ? s : best;
    }
    score[(size_t)y * w + x] = (uint8_t)best;
}

__global__ void kV2(const uint8_t *__restrict__ pad, int w, int h, int pitch, uint8_t *__restrict__ score)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const uint8_t *c = pad + (size_t)(y + 19) * pitch + x + 19;
    const int v = c[0];
    const int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
    const int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
    int d[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) d[k] = v - (int)c[dy[k] * pitch + dx[k]];
    int best = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        int mn = d[k], mx = d[k];
#pragma unroll
        for (int j = 1; j < 9; ++j) { int t = d[(k + j) & 15]; mn = min(mn, t); mx = max(mx, t); }
        asm volatile("" : "+r"(mn), "+r"(mx));
        best = max(best, max(mn, -mx));
    }
    score[(size_t)y * w + x] = (uint8_t)best;
}

int main(){
  int w=64,h=48,pitch=w+38; size_t np=(size_t)pitch*(h+38);
  uint8_t *hp=(uint8_t*)malloc(np); srand(1); for(size_t i=0;i<np;i++) hp[i]=rand()&255;
  uint8_t *dp,*ds; cudaMalloc(&dp,np); cudaMalloc(&ds,w*h); cudaMemcpy(dp,hp,np,cudaMemcpyHostToDevice);
  k<<<dim3(2,6),dim3(32,8)>>>(dp,w,h,pitch,ds);
  uint8_t *hs=(uint8_t*)malloc(w*h); cudaMemcpy(hs,ds,w*h,cudaMemcpyDeviceToHost);
  const int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
  const int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
  int bad=0;
  for(int y=0;y<h;y++)for(int x=0;x<w;x++){ const uint8_t*c=hp+(size_t)(y+19)*pitch+x+19; int d[16]; for(int k=0;k<16;k++) d[k]=c[0]-c[dy[k]*pitch+dx[k]];
    int best=0; for(int k=0;k<16;k++){int mn=d[k],mx=d[k]; for(int j=1;j<9;j++){int t=d[(k+j)&15]; mn=std::min(mn,t); mx=std::max(mx,t);} best=std::max(best,std::max(mn,-mx));}
    if((uint8_t)best!=hs[y*w+x]) bad++; }
  printf("V-orig bad %d of %d (%s)\n",bad,w*h,cudaGetErrorString(cudaGetLastError()));
  void (*ks[3])(const uint8_t*,int,int,int,uint8_t*)={kV0,kV1,kV2};
  for(int vi=0;vi<3;vi++){ ks[vi]<<<dim3(2,6),dim3(32,8)>>>(dp,w,h,pitch,ds); cudaMemcpy(hs,ds,w*h,cudaMemcpyDeviceToHost); int b2=0;
   for(int y=0;y<h;y++)for(int x=0;x<w;x++){ const uint8_t*c=hp+(size_t)(y+19)*pitch+x+19; int d[16]; for(int k=0;k<16;k++) d[k]=c[0]-c[dy[k]*pitch+dx[k]];
    int best=0; for(int k=0;k<16;k++){int mn=d[k],mx=d[k]; for(int j=1;j<9;j++){int t=d[(k+j)&15]; mn=std::min(mn,t); mx=std::max(mx,t);} best=std::max(best,std::max(mn,-mx));}
    if((uint8_t)best!=hs[y*w+x]) b2++; }
   printf("V%d bad %d\n",vi,b2);}
}
