/* TEST INFRASTRUCTURE ONLY.  Probe used to pin the float operation sequences of the reference's SCALAR
 * epilogue paths (they depend on how GCC contracts a*b+c).  It calls the compiled,
 * unmodified reference (oracle/_ref/libcphnsw_refshim.so: refshim_convert with count = 1, so lane 0 takes
 * the scalar tail) on 20 000 random inputs and reports which candidate sequence reproduces every result
 * bit for bit.  Findings (g++ 13.3 -O3 -march=x86-64-v3 -mfma, identical for D = 16..2048, B = 1,2,4):
 *   1-bit tail, N-bit tail estimate, convert_msb_to_lower_bounds:  t = A*fs; t = fma(pc,B,t); t += C
 *   N-bit tail, plane-0 (lower-bound) chain:                       t = B*pc; t = fma(A,fs,t); t += C
 *   affine: fma(a,t,b);  lower: fma(-((nop+nop)*sqrt_dqp), cu, fma(nop,nop,dqp))
 * build: gcc -O1 -ffp-contract=off -o probe probe_contraction.c -ldl -lm
 * run:   ./probe oracle/_ref/libcphnsw_refshim.so <B> <which: 0 nbit-tail lower, 1 msb, 2 nbit-tail est, 3 1-bit tail> <D>
 */
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <dlfcn.h>
typedef int (*conv_t)(uint32_t,uint32_t,const float*,const uint32_t*,const uint32_t*,const uint32_t*,const float*,const float*,const float*,const uint16_t*,const uint16_t*,uint32_t,float,float*,float*,float*);
static float ipa(int v, float A, float fs, float B, float pc, float C){
  switch(v){
   case 0: { float t=A*fs; t=fmaf(pc,B,t); return t+C; }
   case 1: { float t=B*pc; t=fmaf(A,fs,t); return t+C; }
   case 2: return fmaf(A,fs,fmaf(B,pc,C));
   case 3: return (A*fs+B*pc)+C;
   case 4: { float t=fmaf(B,pc,C); return fmaf(A,fs,t);} 
   case 5: { float t=A*fs; float u=B*pc; return (t+u)+C; }
   case 6: { float t=fmaf(A,fs,C); return fmaf(B,pc,t);} 
   case 7: { float u=B*pc; float t=A*fs+u; return t+C; }
  } return 0; }
static float aff(int v, float a, float t, float b){ return v? a*t+b : fmaf(a,t,b); }
static float low(int v, float nop, float dqp, float sq, float cu){
  switch(v){
   case 0: return fmaf(-((nop+nop)*sq), cu, fmaf(nop,nop,dqp));
   case 1: return (nop*nop+dqp) - ((2.0f*nop)*sq)*cu;
   case 2: { float p=((2.0f*nop)*sq)*cu; return fmaf(nop,nop,dqp) - p; }
   case 3: { float p=((2.0f*nop)*sq)*cu; return fmaf(nop,nop,dqp-p); }
   case 4: { float s=nop*nop+dqp; return fmaf(-((2.0f*nop)*sq),cu,s);} 
   case 5: { float s=fmaf(nop,nop,dqp); float p=(2.0f*nop)*sq; return s - p*cu; }
  } return 0; }
int main(int argc,char**argv){
  void* h=dlopen(argv[1],RTLD_NOW); conv_t conv=(conv_t)dlsym(h,"refshim_convert");
  int B=atoi(argv[2]); int which=atoi(argv[3]); int DD=atoi(argv[4]); /* 0: nbit tail lower, 1: msb_lower, 2: tail est, 3: 1bit tail */
  srand(1); int NV=8*2*6; long match[8][2][6]; memset(match,0,sizeof match); long total=0;
  for(int it=0; it<20000; ++it){
    float params[7]; uint32_t nbit[32]={0},msb[32]={0},msb2[32]={0}; float nop[32]={0},ipqo[32]={0},ipcp[32]={0}; uint16_t pop[32]={0},wpop[32]={0};
    #define R ((float)rand()/RAND_MAX)
    params[0]=0.001f+0.01f*R; params[1]=-0.3f*R; params[2]=-1.0f+2*R; params[3]=0.9f+0.2f*R; params[4]=0.05f*(R-0.5f); params[5]=0.3f; params[6]=0.7f+0.3f*R;
    msb[0]=rand()%1900; msb2[0]=rand()%5000; nbit[0]=rand()%(1900*((1<<B)-1)+1); nop[0]=0.4f+2*R; ipqo[0]=0.31f+0.8f*R; ipcp[0]=0.1f*(R-0.5f); pop[0]=rand()%128; wpop[0]=rand()%(128*((1<<B)-1)+1);
    float dqp=1.0f+300*R; float est[32],lower[32],ml[32];
    conv(DD,B,params,nbit,msb,msb2,nop,ipqo,ipcp,pop,wpop,1,dqp,est,lower,ml);
    float K=(float)((1<<B)-1), invK=1.0f/K; float sq=sqrtf(dqp); float q=ipqo[0]>params[5]?ipqo[0]:params[5];
    float target; float A,Bc,fs,pc;
    if(which==0){ target=lower[0]; A=params[0]; Bc=params[1]; fs=(float)msb[0]; pc=(float)pop[0]; }
    else if(which==1){ target=ml[0]; float ik=1.0f/3.0f; A=params[0]*ik; Bc=params[1]*ik; fs=(float)msb2[0]; pc=(float)pop[0]; }
    else if(which==3){ target=lower[0]; A=params[0]; Bc=params[1]; fs=(float)nbit[0]; pc=(float)pop[0]; }
    else { target=est[0]; A=params[0]*invK; Bc=params[1]*invK; fs=(float)nbit[0]; pc=(float)wpop[0]; }
    total++;
    for(int i=0;i<8;i++)for(int j=0;j<2;j++)for(int k=0;k<6;k++){
      float ip=ipa(i,A,fs,Bc,pc,params[2]); float t=(ip-ipcp[0])/q; float e=aff(j,params[3],t,params[4]); float r;
      if(which==2){ float d = (k==0)? fmaf(-(nop[0]+nop[0]),e,fmaf(nop[0],nop[0],dqp)) : (k==1)? (nop[0]*nop[0]+dqp)-(2.0f*nop[0])*e : (k==2)? fmaf(nop[0],nop[0],dqp)-(2.0f*nop[0])*e : (k==3)? fmaf(nop[0],nop[0],dqp-(2.0f*nop[0])*e) : (k==4)? fmaf(-(2.0f*nop[0]),e,nop[0]*nop[0]+dqp): 0; r=d<0?0:d; }
      else { float cu=(e+params[6])/sq; if(cu<-1)cu=-1; if(cu>1)cu=1; r=low(k,nop[0],dqp,sq,cu); if(r<0)r=0; }
      if(memcmp(&r,&target,4)==0) match[i][j][k]++;
    }
  }
  for(int i=0;i<8;i++)for(int j=0;j<2;j++)for(int k=0;k<6;k++) if(match[i][j][k]==total) printf("ipa=%d aff=%d low=%d : %ld/%ld\n",i,j,k,match[i][j][k],total);
  return 0; }
