// Simulation behind dsp_dcwin.h: how far the float DC-blocker chain strays from the exact (double) state at the 1024-sample tile
// boundaries, which boundaries are enumerable ("good": exact state within `slack` binades of the largest state four boundaries back and
// one ahead) and how often the true state leaves a window of +-`half` float steps there.  gcc -O1 dc_window_sim.c -lm; ./a.out <mode> <half> <slack>
// modes: 0 noise + DC offset, 1 zero-mean noise, 2 carrier next to 0 Hz, 3 strong carrier + offset, 4 slow carrier, 5 FM sweep.
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <string.h>
#include <stdint.h>
static double rnd(void){ return (rand()+0.5)/(RAND_MAX+1.0); }
static double gauss(void){ return sqrt(-2*log(rnd()))*cos(6.283185307179586*rnd()); }
static int ordi(float f){ int32_t i; memcpy(&i,&f,4); return i>=0? i : (int)(0x80000000u-(uint32_t)i); }
static int expo(float f){ int e; frexpf(f,&e); return f==0? -200: e; }
int main(int argc,char**argv){
  int mode=argc>1?atoi(argv[1]):0; int half=argc>2?atoi(argv[2]):128; int slack=argc>3?atoi(argv[3]):1;
  long N=40000000; const float c=0.9995f; const double a=(double)c;
  long T=N/1024; float*ef=malloc(sizeof(float)*(T+2)); float*vf=malloc(sizeof(float)*(T+2));
  volatile float v=0; double e=0;
  for(long n=0;n<N;n++){
    if((n&1023)==0){ ef[n>>10]=(float)e; vf[n>>10]=v; }
    double xr;
    if(mode==0) xr=0.05*gauss()+0.01;
    else if(mode==1) xr=0.05*gauss();
    else if(mode==2) xr=0.02*cos(0.0056*n)+0.007*gauss();
    else if(mode==3) xr=0.3*cos(0.3*n)+0.02*cos(0.0056*n)+0.01+0.007*gauss();
    else if(mode==4) xr=0.02*cos(0.00056*n)+0.007*gauss();
    else xr=0.2*cos(0.05*n+3*sin(1e-4*n))+0.001*gauss();
    float x=(float)xr;
    volatile float t=c*v; volatile float nv=x+t; v=nv;
    e=a*e+(double)x;
  }
  long good=0,miss=0,tot=0,run=0,maxrun=0; long runs[8]={0};
  for(long b=50;b<T-2;b++){
    float M=0; for(long k=b-4;k<=b+1;k++){ float m=fabsf(ef[k]); if(m>M)M=m; }
    int g = expo(ef[b]) >= expo(M)-slack;
    tot++;
    if(g){ good++; long d=labs((long)ordi(vf[b])-(long)ordi(ef[b])); if(d>=half)miss++; if(run>maxrun)maxrun=run; runs[run<7?run:7]++; run=0; } else run++;
  }
  printf("mode %d half %d slack %d: good %.3f, misses %ld of %ld good, max bad run %ld; run hist:",mode,half,slack,(double)good/tot,miss,good,maxrun);
  for(int i=0;i<8;i++)printf(" %ld",runs[i]); printf("\n");
  return 0;
}
