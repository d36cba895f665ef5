/* synth.c — deterministic synthetic LAS clouds for the five BASELINE.json configs.
 *
 * Integer-only arithmetic (no libm), counter-based hashing: the same (scene, seed, region)
 * yields the same bytes on every machine and for every thread count.  Modelled on the
 * reference's testpattern.cpp (laserize(): returnNum=nReturns=1, intensity 1024,
 * gpsTime = point index, testpattern.cpp:166-190), but sampled on a jittered grid so
 * that no two points share XY (the reference silently merges exact-duplicate XYZ,
 * octree.cpp:626-644, which would make point counts incomparable).
 *
 * Scenes:
 *   1  "street": gently crowned ground + 12x12 m boxes / gabled roofs every 40 m   (C1, 20 pts/m2)
 *   2  "aerial": sum of 3 sinusoids + N(0,0.02 m) noise + 15 % vegetation U(0,8 m) (C2/C3, 20 pts/m2)
 *   4  "terrestrial": 4 scanner positions, density quadrupling towards each scanner,
 *                     walls and poles, 15 m relief, 0.1 mm scale                   (C4)
 *   5  "urban": 45-degree sawtooth hillside, flat roofs, walls, overhangs          (C5)
 */
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include "wb_synth.h"

static inline uint64_t mix64(uint64_t z)
{
  z+=0x9e3779b97f4a7c15ull;
  z=(z^(z>>30))*0xbf58476d1ce4e5b9ull;
  z=(z^(z>>27))*0x94d049bb133111ebull;
  return z^(z>>31);
}

static inline int32_t isin_q15(uint32_t ph)
/* sine of a 32-bit binary angle, result in [-32767,32767]; parabolic + one correction. */
{
  int64_t x=(int32_t)ph>>8;               /* [-2^23,2^23) = [-pi,pi) */
  int64_t ax=x<0?-x:x;
  int64_t y=(x*((1<<23)-ax))>>29;         /* 4x(1-|x|) in Q15 */
  int64_t ay=y<0?-y:y;
  y+=(225*(((y*ay)>>15)-y))/1000;
  if (y>32767) y=32767;
  if (y<-32767) y=-32767;
  return (int32_t)y;
}

static inline int64_t wave_mm(int64_t pos_mm,int64_t wavelength_mm,int64_t amp_mm,uint32_t phase0)
{
  uint64_t u=(uint64_t)(pos_mm%wavelength_mm+wavelength_mm)%(uint64_t)wavelength_mm;
  uint32_t ph=(uint32_t)((u<<32)/(uint64_t)wavelength_mm)+phase0;
  return (amp_mm*isin_q15(ph))/32767;
}

static inline int64_t gauss_mm(uint64_t h,int64_t sigma_mm)
/* Irwin-Hall(4) approximation of N(0,sigma): sum of four 16-bit uniforms. */
{
  int64_t s=(int64_t)(h&0xffff)+(int64_t)((h>>16)&0xffff)+(int64_t)((h>>32)&0xffff)+(int64_t)((h>>48)&0xffff)-2*65535;
  /* variance of the sum = 4*65536^2/12 -> std = 37837 */
  return (s*sigma_mm)/37837;
}

static inline int64_t tri(int64_t pos,int64_t period)
/* triangle wave in [0,period/2] */
{
  int64_t u=((pos%period)+period)%period;
  return u<period/2?u:period-u;
}

/* ---- scene surfaces: (x,y) in 0.1 mm units relative to the scene origin -> z in the same units */

static int64_t scene_street(int64_t x,int64_t y,uint64_t h)
{
  const int64_t M=10000; /* 1 m */
  int64_t z=100*M+wave_mm(x,300*M,15*M/10,0)+wave_mm(y,220*M,1*M,0x40000000u);
  int64_t road=tri(x,120*M);                      /* street every 120 m, 15 m wide, crowned */
  if (road<72*M/10)
    z-=road/50;
  else if (road<74*M/10)
    z-=15*M/100;
  int64_t cx=x/(40*M),cy=y/(40*M);
  int64_t lx=x-cx*40*M-20*M,ly=y-cy*40*M-20*M;  /* local coords about the cell centre */
  int64_t alx=lx<0?-lx:lx,aly=ly<0?-ly:ly;
  if (alx<6*M && aly<6*M && road>=10*M)
  {
    int64_t base=100*M+wave_mm(cx*40*M+20*M,300*M,15*M/10,0)+wave_mm(cy*40*M+20*M,220*M,1*M,0x40000000u);
    if ((cx+cy)&1)
      z=base+4*M+(6*M-aly)/3;                    /* gabled roof, ridge along x */
    else
      z=base+6*M;                                 /* flat box */
  }
  return z+gauss_mm(h,M/100);
}

static int64_t scene_aerial(int64_t x,int64_t y,uint64_t h)
{
  const int64_t M=10000;
  int64_t z=120*M+wave_mm(x,107*M,3*M,0)+wave_mm(y,145*M,2*M,0x20000000u)+wave_mm(x+y,61*M,1*M,0x90000000u);
  uint64_t h2=mix64(h);
  z+=gauss_mm(h,2*M/100);
  if (h2%100<15)
    z+=(int64_t)((h2>>8)%(uint64_t)(8*M));
  return z;
}

static int64_t scene_urban(int64_t x,int64_t y,uint64_t h)
{
  const int64_t M=10000;
  uint64_t h2=mix64(h);
  int64_t z=50*M+tri(x,400*M)*95/100+wave_mm(y,180*M,2*M,0);  /* ~43.5 degree sawtooth */
  int64_t cx=x/(50*M),cy=y/(50*M);
  int64_t lx=x-cx*50*M-25*M,ly=y-cy*50*M-25*M;
  int64_t alx=lx<0?-lx:lx,aly=ly<0?-ly:ly;
  int64_t m=alx>aly?alx:aly;
  if (m<10*M)
  {
    int64_t base=50*M+tri(cx*50*M+25*M,400*M)*95/100+wave_mm(cy*50*M+25*M,180*M,2*M,0);
    int64_t roof=base+12*M;
    if (m<8*M)
      z=roof;                                     /* flat roof */
    else if (m<82*M/10)
      z=z+(int64_t)((h2>>8)%(uint64_t)(roof-z>M?roof-z:M)); /* wall: stacked in z */
    else if ((h2&1) && roof-3*M>z)
      z=roof-3*M;                                 /* overhang slab; the other half stay ground */
  }
  return z+gauss_mm(h,15*M/1000);
}

static int64_t scene_terr_surface(int64_t x,int64_t y,uint64_t h)
{
  const int64_t M=10000;
  uint64_t h2=mix64(h);
  int64_t z=20*M+wave_mm(x,97*M,5*M,0)+wave_mm(y,71*M,25*M/10,0x30000000u);
  int64_t px=tri(x,20*M),py=tri(y,20*M);
  if (px<15*M/100 && py<15*M/100)
    z+=(int64_t)((h2>>8)%(uint64_t)(8*M));       /* pole */
  else if (tri(x+5*M,60*M)<10*M/100 && tri(y,60*M)>5*M)
    z+=(int64_t)((h2>>8)%(uint64_t)(5*M));       /* wall */
  return z+gauss_mm(h,3*M/1000);
}

/* ---- record writer ------------------------------------------------------ */

static const int kRecLen[11]={20,28,26,34,57,63,30,36,38,59,67}; /* LAS spec; las.cpp:38 */

int wb_synth_record_length(int fmt)
{
  return (fmt<0 || fmt>10)?0:kRecLen[fmt];
}

static inline void put_record(uint8_t *r,int fmt,int32_t xi,int32_t yi,int32_t zi,double gps,uint64_t h)
{
  int len=kRecLen[fmt],o;
  memset(r,0,len);
  memcpy(r,&xi,4);
  memcpy(r+4,&yi,4);
  memcpy(r+8,&zi,4);
  r[12]=0x00; r[13]=0x04;                         /* intensity 1024 */
  if (fmt<6)
  {
    r[14]=0x09;                                   /* return 1 of 1 */
    o=20;
  }
  else
  {
    r[14]=0x11;
    o=22;
  }
  if (fmt==1 || fmt>=3)
  {
    memcpy(r+o,&gps,8);
    o+=8;
  }
  if (fmt==2 || fmt==3 || fmt==5 || fmt==7 || fmt==8 || fmt==10)
  {
    uint16_t c=(uint16_t)(30000+(h&0xfff));
    memcpy(r+o,&c,2); memcpy(r+o+2,&c,2); memcpy(r+o+4,&c,2);
  }
}

/* ---- public API --------------------------------------------------------- */

int wb_synth_describe(int scene,uint64_t n_points,wb_synth_desc *d)
{
  /* Units: the jittered grid is laid out in integer "ticks" (scale metres per tick). */
  double density;
  memset(d,0,sizeof(*d));
  d->scene=scene;
  switch (scene)
  {
    case 1: density=20; d->scale=0.001; d->fmt=1; break;
    case 2: density=20; d->scale=0.001; d->fmt=1; break;
    case 3: density=20; d->scale=0.001; d->fmt=6; scene=2; break;   /* C3 = aerial model in fmt 6 */
    case 4: density=2000; d->scale=0.0001; d->fmt=1; break;
    case 5: density=20; d->scale=0.001; d->fmt=3; break;
    default: return -1;
  }
  d->offset[0]=500000; d->offset[1]=4200000; d->offset[2]=0;
  if (scene==4)
  {
    /* 4 scanners x 6 nested levels; each (scanner,level) annulus gets an equal share. */
    uint64_t per=n_points/24,g=1;
    while ((g+1)*(g+1)<=per) g++;
    d->grid_nx=d->grid_ny=g;                       /* g x g cells per (scanner,level) */
    d->n_points=g*g*24;
    d->extent_ticks=1580000;                       /* 158 m in 0.1 mm ticks */
    d->cell_ticks=0;
    return 0;
  }
  {
    uint64_t g=1;
    while ((g+1)*(g+1)<=n_points) g++;
    /* cell edge in ticks so that density is ~20 pts/m2: 1/sqrt(20) m = 223.6 mm */
    d->cell_ticks=224; (void)density;
    d->grid_nx=d->grid_ny=g;
    d->n_points=g*g;
    d->extent_ticks=g*d->cell_ticks;
  }
  return 0;
}

static uint64_t perm_multiplier(uint64_t m)
/* a multiplier near m/phi that is coprime to m, so i -> (i*p+q) mod m is a bijection on [0,m). */
{
  uint64_t p;
  if (m<3)
    return 1;
  p=0x9e3779b97f4a7c15ull%m;
  if (p<2)
    p=2;
  while (1)
  {
    uint64_t a=p,b=m;
    while (b) { uint64_t t=a%b; a=b; b=t; }
    if (a==1)
      return p;
    p++;
    if (p>=m)
      p=2;
  }
}

static inline uint64_t perm_index(uint64_t i,uint64_t m,uint64_t p,uint64_t seed)
{
  return (uint64_t)(((unsigned __int128)i*p+seed%m)%m);
}

int wb_synth_generate(const wb_synth_desc *d,uint64_t seed,
                      uint64_t cell_x0,uint64_t cell_y0,uint64_t ncx,uint64_t ncy,
                      uint64_t gps_base,uint8_t *recs,int32_t bbox[6])
/* Fills ncx*ncy records for the grid sub-rectangle [cell_x0,cell_x0+ncx) x [cell_y0,...).
 * Record k holds grid cell perm(k) of the sub-rectangle; gpsTime = gps_base + k.
 * bbox = {minx,miny,minz,maxx,maxy,maxz} in record integer units. */
{
  int fmt=d->fmt,len=kRecLen[fmt],scene=d->scene==3?2:d->scene;
  uint64_t m=ncx*ncy,pm=perm_multiplier(ncx*ncy);
  int32_t bb[6]={INT32_MAX,INT32_MAX,INT32_MAX,INT32_MIN,INT32_MIN,INT32_MIN};
  if (scene==4)
    return -2;
  int64_t unit=d->scale==0.001?10:1;               /* scene functions work in 0.1 mm */
  long long k;
  #pragma omp parallel
  {
    int32_t lb[6]={INT32_MAX,INT32_MAX,INT32_MAX,INT32_MIN,INT32_MIN,INT32_MIN};
    #pragma omp for schedule(static)
    for (k=0;k<(long long)m;k++)
    {
      uint64_t c=perm_index((uint64_t)k,m,pm,seed);
      uint64_t cx=cell_x0+c%ncx,cy=cell_y0+c/ncx;
      uint64_t h=mix64(seed*0x100000001b3ull+cy*d->grid_nx+cx);
      int64_t x=(int64_t)(cx*d->cell_ticks+(h>>40)%d->cell_ticks);
      int64_t y=(int64_t)(cy*d->cell_ticks+(h>>20&0xfffff)%d->cell_ticks);
      uint64_t h3=mix64(h^0xabcdef);
      int64_t z;
      switch (scene)
      {
        case 1: z=scene_street(x*unit,y*unit,h3); break;
        case 5: z=scene_urban(x*unit,y*unit,h3); break;
        default: z=scene_aerial(x*unit,y*unit,h3);
      }
      z/=unit;
      int32_t xi=(int32_t)x,yi=(int32_t)y,zi=(int32_t)z;
      put_record(recs+(size_t)k*len,fmt,xi,yi,zi,(double)(gps_base+(uint64_t)k),h3);
      if (xi<lb[0]) lb[0]=xi; if (yi<lb[1]) lb[1]=yi; if (zi<lb[2]) lb[2]=zi;
      if (xi>lb[3]) lb[3]=xi; if (yi>lb[4]) lb[4]=yi; if (zi>lb[5]) lb[5]=zi;
    }
    #pragma omp critical
    {
      int q;
      for (q=0;q<3;q++) if (lb[q]<bb[q]) bb[q]=lb[q];
      for (q=3;q<6;q++) if (lb[q]>bb[q]) bb[q]=lb[q];
    }
  }
  if (bbox)
    memcpy(bbox,bb,sizeof(bb));
  return 0;
}

int wb_synth_generate_terrestrial(const wb_synth_desc *d,uint64_t seed,uint64_t gps_base,
                                  uint8_t *recs,int32_t bbox[6])
/* Scene 4: for scanner s (0..3) and level l (0..5) a g x g jittered grid over the square of
 * half-size 79 m / 2^l centred on the scanner, clipped to the 158 m scene by wrapping.
 * Points of different (s,l) live on disjoint residue classes (x mod 5, y mod 5), so no two
 * points share XY.  Density roughly quadruples per level, i.e. ~1/r^2. */
{
  static const int64_t sx[4]={400000,1180000,400000,1180000},sy[4]={400000,400000,1180000,1180000};
  int fmt=d->fmt,len=kRecLen[fmt];
  uint64_t g=d->grid_nx,per=g*g,m=per*24,pm=perm_multiplier(per*24);
  int32_t bb[6]={INT32_MAX,INT32_MAX,INT32_MAX,INT32_MIN,INT32_MIN,INT32_MIN};
  long long k;
  #pragma omp parallel
  {
    int32_t lb[6]={INT32_MAX,INT32_MAX,INT32_MAX,INT32_MIN,INT32_MIN,INT32_MIN};
    #pragma omp for schedule(static)
    for (k=0;k<(long long)m;k++)
    {
      uint64_t c=perm_index((uint64_t)k,m,pm,seed);
      uint64_t sl=c/per,ci=c%per;
      int s=(int)(sl&3),l=(int)(sl>>2),ra=(int)(sl%5),rb=(int)(sl/5);
      int64_t half=790000>>l;
      int64_t cell=2*half/(int64_t)g;               /* >= 5 ticks by construction of g */
      uint64_t h=mix64(seed*0x100000001b3ull+c);
      int64_t x0=sx[s]-half+(int64_t)(ci%g)*cell,y0=sy[s]-half+(int64_t)(ci/g)*cell;
      if (cell<5) cell=5;
      int64_t nx=(cell-1-(((ra-x0)%5+5)%5))/5+1,ny=(cell-1-(((rb-y0)%5+5)%5))/5+1;
      int64_t x=x0+(((ra-x0)%5+5)%5)+5*(int64_t)((h>>40)%(uint64_t)(nx>0?nx:1));
      int64_t y=y0+(((rb-y0)%5+5)%5)+5*(int64_t)((h>>20&0xfffff)%(uint64_t)(ny>0?ny:1));
      x=((x%1580000)+1580000)%1580000;              /* wrap preserves the residue: 1580000 % 5 == 0 */
      y=((y%1580000)+1580000)%1580000;
      uint64_t h3=mix64(h^0xabcdef);
      int64_t z=scene_terr_surface(x,y,h3);
      int32_t xi=(int32_t)x,yi=(int32_t)y,zi=(int32_t)z;
      put_record(recs+(size_t)k*len,fmt,xi,yi,zi,(double)(gps_base+(uint64_t)k),h3);
      if (xi<lb[0]) lb[0]=xi; if (yi<lb[1]) lb[1]=yi; if (zi<lb[2]) lb[2]=zi;
      if (xi>lb[3]) lb[3]=xi; if (yi>lb[4]) lb[4]=yi; if (zi>lb[5]) lb[5]=zi;
    }
    #pragma omp critical
    {
      int q;
      for (q=0;q<3;q++) if (lb[q]<bb[q]) bb[q]=lb[q];
      for (q=3;q<6;q++) if (lb[q]>bb[q]) bb[q]=lb[q];
    }
  }
  if (bbox)
    memcpy(bbox,bb,sizeof(bb));
  return 0;
}

static void put16(uint8_t *p,uint16_t v) { memcpy(p,&v,2); }
static void put32(uint8_t *p,uint32_t v) { memcpy(p,&v,4); }
static void put64(uint8_t *p,uint64_t v) { memcpy(p,&v,8); }
static void putd(uint8_t *p,double v) { memcpy(p,&v,8); }

int wb_synth_header(const wb_synth_desc *d,uint64_t n_points,const int32_t bbox[6],uint8_t *hdr)
/* Writes a LAS 1.2 (formats 0-3, 227 bytes) or LAS 1.4 (formats 6-10, 375 bytes) public
 * header block laid out as LasHeader::openRead parses it (las.cpp:299-428).  Returns its size. */
{
  int v14=d->fmt>=6,size=v14?375:227,i;
  memset(hdr,0,size);
  memcpy(hdr,"LASF",4);
  hdr[24]=1; hdr[25]=v14?4:2;
  memcpy(hdr+26,"wolkenbase_b200 synth",21);
  memcpy(hdr+58,"wb_synth",8);
  put16(hdr+90,1); put16(hdr+92,2026);
  put16(hdr+94,(uint16_t)size);
  put32(hdr+96,(uint32_t)size);
  put32(hdr+100,0);
  hdr[104]=(uint8_t)d->fmt;
  put16(hdr+105,(uint16_t)kRecLen[d->fmt]);
  if (!v14)
  {
    put32(hdr+107,(uint32_t)n_points);
    put32(hdr+111,(uint32_t)n_points);             /* all points are return 1 */
  }
  for (i=0;i<3;i++)
  {
    putd(hdr+131+8*i,d->scale);
    putd(hdr+155+8*i,d->offset[i]);
  }
  /* max/min pairs, computed exactly as readPoint computes coordinates (las.cpp:808) */
  for (i=0;i<3;i++)
  {
    double mx=d->offset[i]+d->scale*bbox[3+i],mn=d->offset[i]+d->scale*bbox[i];
    putd(hdr+179+16*i,mx);
    putd(hdr+187+16*i,mn);
  }
  if (v14)
  {
    put64(hdr+247,n_points);
    put64(hdr+255,n_points);
  }
  return size;
}
