// Model of k_lz_find's chain walks on the CPU: hop statistics of the bucket-chain walk with and without SKIP LINKS
// (link2[i] = most recent earlier node of the bucket whose trigram differs from trigram(i)), chunk = 256 KiB, 14-bit hash.
//   python -c "from libflate_b200 import titles; open('/tmp/t16.bin','wb').write(bytes(titles.generate(16<<20, seed=42)))"
//   gcc -O2 -o /tmp/chain_hops tools/chain_hops_model.c && /tmp/chain_hops 14 16 /tmp/t16.bin
// Result on titles-shaped text: plain chains 1.08 hops/position, 6.3 % of the positions need more than one hop, 0.32 % more than
// 16 (deferred to k_lz_fixup); skip links built in position order: 4.5 % / 0.003 %.  The kernel upgrades links without ordering
//  and lands in between (0.10 %).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
// hop statistics with "skip links": link2[i] = most recent earlier position of the bucket whose trigram differs from trigram(i)
int main(int argc,char**argv){
  int HB=atoi(argv[1]); int CAP=atoi(argv[2]); const char*fn=argc>3?argv[3]:"/tmp/t16.bin";
  FILE*f=fopen(fn,"rb"); static uint8_t buf[16<<20]; size_t n=fread(buf,1,sizeof buf,f);
  uint32_t *head=malloc(4u<<HB); static uint32_t link2[1<<18];
  double sum_hops=0; long npos=0,ndef=0, hist[40]={0}; long found=0; double sum_max=0; long nstep=0; double q_sum=0;
  for(size_t c0=0;c0+262144<=n;c0+=262144){
    uint8_t*p=buf+c0; memset(head,0,4u<<HB);
    static int hop[262144];
    for(uint32_t i=0;i<262144-3;i++){
      uint32_t t=p[i]|(p[i+1]<<8)|(p[i+2]<<16); uint32_t h=(t*0x9E3779B1u)>>(32-HB);
      uint32_t o=head[h]; head[h]=i+1; uint32_t d=o?i+1-o:0; if(d>32768)d=0;
      // link2
      uint32_t l2=0;
      if(d){ uint32_t c=i-d; uint32_t tc=p[c]|(p[c+1]<<8)|(p[c+2]<<16); if(tc!=t) l2=d; else { l2 = link2[c]? d+link2[c]:0; } if(l2>32768) l2=0; }
      link2[i]=l2;
      // walk
      uint32_t total=0,j=i,hops=0; int res=0; uint32_t dd=d;
      while(dd){ total+=dd; if(total>32768)break; j-=dd; hops++; uint32_t tj=p[j]|(p[j+1]<<8)|(p[j+2]<<16); if(tj==t){res=1;break;} dd=link2[j]; if(hops==CAP){ if(dd)res=2; break;} }
      hop[i]=hops; sum_hops+=hops; npos++; if(res==2)ndef++; if(res==1)found++; hist[hops<39?hops:39]++;
    }
    for(uint32_t b=0;b+32<=262144-3;b+=32){ int mx=0,q=0; for(int l=0;l<32;l++){int h=hop[b+l]; if(h>mx)mx=h; if(h>=2)q++;} sum_max+=mx; nstep++; q_sum+=q; }
  }
  printf("HB=%d cap=%d mean hops %.3f  mean max/step %.2f  queued/step %.2f deferred %.4f%% found %.1f%%\n",HB,CAP,sum_hops/npos,sum_max/nstep,q_sum/nstep,100.0*ndef/npos,100.0*found/npos);
  for(int i=0;i<=17;i++)printf("%d:%.3f ",i,100.0*hist[i]/npos); printf("\n");
}
