#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__device__ __noinline__ void mic_general(const double *__restrict__ ortho, const double *__restrict__ frac, double &x, double &y,
                                         double &z) {
    double f0 = frac[0] * x + frac[1] * y + frac[2] * z;
    double f1 = frac[3] * x + frac[4] * y + frac[5] * z;
    double f2 = frac[6] * x + frac[7] * y + frac[8] * z;
    f0 -= rint(f0);
    f1 -= rint(f1);
    f2 -= rint(f2);
    const double bx = ortho[0] * f0 + ortho[1] * f1 + ortho[2] * f2;
    const double by = ortho[3] * f0 + ortho[4] * f1 + ortho[5] * f2;
    const double bz = ortho[6] * f0 + ortho[7] * f1 + ortho[8] * f2;
    double best = bx * bx + by * by + bz * bz;
    x = bx;
    y = by;
    z = bz;
    for (int h = -1; h <= 1; ++h)
        for (int k = -1; k <= 1; ++k)
            for (int l = -1; l <= 1; ++l) {
                if (h == 0 && k == 0 && l == 0) continue;
                const double cx = bx + ortho[0] * h + ortho[1] * k + ortho[2] * l;
                const double cy = by + ortho[3] * h + ortho[4] * k + ortho[5] * l;
                const double cz = bz + ortho[6] * h + ortho[7] * k + ortho[8] * l;
                const double d2 = cx * cx + cy * cy + cz * cz;
                if (d2 < best) {
                    best = d2;
                    x = cx;
                    y = cy;
                    z = cz;
                }
            }
}


__global__ void k(const double* o, const double* f, double* v, int n){ int i=blockIdx.x*blockDim.x+threadIdx.x; if(i<n){ double x=v[3*i],y=v[3*i+1],z=v[3*i+2]; mic_general(o,f,x,y,z); v[3*i]=x;v[3*i+1]=y;v[3*i+2]=z; } }
int main(){ double L=24.66; double cell[9]={L,0,0,3.0,L,0,-2.0,4.0,L}; double o[9],f[9];
 for(int r=0;r<3;r++)for(int c=0;c<3;c++) o[r*3+c]=cell[c*3+r];
 // inverse of o
 double det=o[0]*(o[4]*o[8]-o[5]*o[7])-o[1]*(o[3]*o[8]-o[5]*o[6])+o[2]*(o[3]*o[7]-o[4]*o[6]);
 f[0]=(o[4]*o[8]-o[5]*o[7])/det; f[1]=(o[2]*o[7]-o[1]*o[8])/det; f[2]=(o[1]*o[5]-o[2]*o[4])/det;
 f[3]=(o[5]*o[6]-o[3]*o[8])/det; f[4]=(o[0]*o[8]-o[2]*o[6])/det; f[5]=(o[2]*o[3]-o[0]*o[5])/det;
 f[6]=(o[3]*o[7]-o[4]*o[6])/det; f[7]=(o[1]*o[6]-o[0]*o[7])/det; f[8]=(o[0]*o[4]-o[1]*o[3])/det;
 const int n=4; double v[12]={30,-20,10, 1,2,3, -13,12.5,-12, 24,24,24};
 double *dv,*dd,*df; cudaMalloc(&dv,sizeof(v)); cudaMalloc(&dd,72); cudaMalloc(&df,72);
 cudaMemcpy(dv,v,sizeof(v),cudaMemcpyHostToDevice); cudaMemcpy(dd,o,72,cudaMemcpyHostToDevice); cudaMemcpy(df,f,72,cudaMemcpyHostToDevice);
 k<<<1,32>>>(dd,df,dv,n); cudaMemcpy(v,dv,sizeof(v),cudaMemcpyDeviceToHost);
 for(int i=0;i<n;i++) printf("%.6f %.6f %.6f  |v|=%.6f\n",v[3*i],v[3*i+1],v[3*i+2],sqrt(v[3*i]*v[3*i]+v[3*i+1]*v[3*i+1]+v[3*i+2]*v[3*i+2]));
 printf("%s\n",cudaGetErrorString(cudaDeviceSynchronize())); }
