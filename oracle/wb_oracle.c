/* wb_oracle.c — CPU restatement of the reference's ground-extraction path (plain C).
 *
 * TEST INFRASTRUCTURE ONLY (see wb_oracle.h).  Each function cites the reference code
 * whose arithmetic it restates; nothing here is copied, the algorithms are re-expressed
 * over flat arrays.  Compile with -ffp-contract=off: the reference is built without FMA.
 *
 * Parity: pinned against oracle/_ref (the compiled reference) by tests/test_oracle_ref.py
 * (run in the build container, artefacts committed under tests/golden/) and against the
 * reference's known-answer tests by tests/test_oracle_kat.py.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include <float.h>
#include "wb_oracle.h"

#define DEG30  0x0aaaaaab   /* angle.h:105 */
#define DEG45  0x10000000
#define DEG90  0x20000000
#define DEG144 0x33333333   /* angle.h:115 */
#define DEG180 0x40000000
#define SQRT_3_4 0.86602540378443864676372317   /* eisenstein.h:33 */
#define SQRT7 2.6457513110645905905016           /* scan.h:23 */

/* ======================= angle.cpp ======================================= */

static double tanTable[511],cosTable[512],sinTable[512];
static int tablesFilled=0;

void wbo_fill_tan_tables(void)
/* angle.cpp:305-320 with sin/cos/tan(int) of angle.cpp:45-63: long-double libm at
 * angle*pi/2^30, rounded to double. */
{
  int i;
  for (i=0;i<511;i++)
    tanTable[i]=(double)tanl((i*0x100000-0xff00000)*M_PIl/1073741824.);
  for (i=0;i<512;i++)
  {
    sinTable[i]=(double)sinl((i*0x100000-0xff80000)*M_PIl/1073741824.);
    cosTable[i]=(double)cosl((i*0x100000-0xff80000)*M_PIl/1073741824.);
  }
  tablesFilled=1;
}

const double *wbo_tan_table(void) { if (!tablesFilled) wbo_fill_tan_tables(); return tanTable; }
const double *wbo_cos_table(void) { if (!tablesFilled) wbo_fill_tan_tables(); return cosTable; }
const double *wbo_sin_table(void) { if (!tablesFilled) wbo_fill_tan_tables(); return sinTable; }

int wbo_atan2i(double y,double x)
/* angle.cpp:117-155: octant folding, 9-step bisection on tanTable, rotation by the bin
 * centre, cubic arctangent correction evaluated in long double (M_PIl makes the whole
 * expression long double), lrint. */
{
  int ret=0,h;
  double t,nx;
  if (!tablesFilled)
    wbo_fill_tan_tables();
  if (x<0)
  {
    ret+=(y>0)?DEG180:-DEG180;
    y=-y;
    x=-x;
  }
  if (y>x)
  {
    ret+=DEG90;
    t=x; x=y; y=-t;
  }
  if (-y>x)
  {
    ret-=DEG90;
    t=x; x=-y; y=t;
  }
  t=y/x;
  for (h=DEG45/2;h>DEG45/1024;h/=2)
    if (t>tanTable[(((ret+DEG45)&0x1ff00000)>>20)-1])
      ret+=h;
    else
      ret-=h;
  h=511-(((ret+DEG45)&0x1ff00000)>>20);
  nx=x*cosTable[h]-y*sinTable[h];
  y=y*cosTable[h]+x*sinTable[h];
  x=nx;
  {
    double q=y/x,c=q*q*q;
    long double v=0x40000000/M_PIl*y/x-1.1392738508503886e8*c;
    ret+=(int)lrintl(v);
  }
  if (x==0 && y==0)
    ret=0;
  return ret;
}

/* ======================= shape.cpp ======================================= */

int wbo_hyperboloid_in(const double v[3],double r,double s,const double p[3])
/* Hyperboloid ctor + in(): shape.cpp:119-135 */
{
  double por=r*(s*s),por2=por*por;
  double cz=v[2]+por;
  double d=hypot(v[0]-p[0],v[1]-p[1]);
  double zd=cz-p[2],ds=d*s;
  if (s>0)
    return zd>0 && zd*zd-ds*ds>=por2;
  return zd<0 && zd*zd-ds*ds>=por2;
}

int wbo_cylinder_in(double cx,double cy,double r,double px,double py)
/* shape.cpp:214-218 */
{
  return hypot(cx-px,cy-py)<=r;
}

int wbo_cylinder_intersects_cube(double cx,double cy,double r,const double c[3],double side)
/* Shape::intersect = in(closestPoint(cube)): shape.cpp:63-66, 220-238 */
{
  double x=c[0],y=c[1];
  if (fabs(cx-x)<side/2) x=cx; else if (cx>x) x+=side/2; else x-=side/2;
  if (fabs(cy-y)<side/2) y=cy; else if (cy>y) y+=side/2; else y-=side/2;
  return wbo_cylinder_in(cx,cy,r,x,y);
}

int wbo_shape_in(int type,const double q[6],const double p[3])
/* Shape::in for every shape of shape.cpp; q = the constructor's arguments.
 * 0 Sphere(c,r) 175-178 | 1 Paraboloid(v,rc) 81-90 | 2 Hyperboloid(v,r,s) 119-135 |
 * 3 Cylinder(c,r) 215-219 | 4 Column(c,side) 252-255.  dist(): point.cpp:189-192, 383-386. */
{
  switch (type)
  {
    case 0:
      return hypot(hypot(p[0]-q[0],p[1]-q[1]),p[2]-q[2])<=q[3];
    case 1:
    {
      double xydist=hypot(q[0]-p[0],q[1]-p[1]),zdist=q[2]-p[2],r=q[3],t;
      if (r==0)
        return xydist==0;
      t=xydist/r;
      return 2*zdist/r>=t*t;
    }
    case 2:
      return wbo_hyperboloid_in(q,q[3],q[4],p);
    case 3:
      return wbo_cylinder_in(q[0],q[1],q[2],p[0],p[1]);
    case 4:
      return fabs(q[0]-p[0])<=q[2]/2 && fabs(q[1]-p[1])<=q[2]/2;
  }
  return 0;
}

int wbo_shape_intersects_cube(int type,const double q[6],const double c[3],double side)
/* Shape::intersect = in(closestPoint(cube)), shape.cpp:63-66 with the closestPoint of each shape
 * (92-117, 137-159, 180-206, 220-238, 257-274) */
{
  double ax=q[0],ay=q[1],x=c[0],y=c[1],z=c[2],pt[3];
  if (fabs(ax-x)<side/2) x=ax; else if (ax>x) x+=side/2; else x-=side/2;
  if (fabs(ay-y)<side/2) y=ay; else if (ay>y) y+=side/2; else y-=side/2;
  if (type==0)
  {
    if (fabs(q[2]-z)<side/2) z=q[2]; else if (q[2]>z) z+=side/2; else z-=side/2;
  }
  else if (type==1 || type==2)
  {
    double sgn=type==1?q[3]:q[4];
    if (sgn>0) z-=side/2;
    if (sgn<0) z+=side/2;
  }
  pt[0]=x; pt[1]=y; pt[2]=z;
  return wbo_shape_in(type,q,pt);
}

void wbo_shape_filter(int type,const double q[6],const double *pts,uint64_t n,uint8_t *in)
{
  uint64_t i;
  for (i=0;i<n;i++)
    in[i]=(uint8_t)wbo_shape_in(type,q,pts+3*i);
}

/* ======================= flowsnake.cpp / eisenstein.cpp =================== */

static const double squareSides[12]=
{ /* flowsnake.cpp:30-44 */
  0.6583539906808145,1.8501627472990723,4.286014912881196,12.6716160058597,
  32.85016274729906,79.01914481623778,243.0343734125204,592.8501627472989,
  1510.850162747299,4399.956311577825,10731.850162747294,29198.8501627473
};
static const int loLim[12]={0,-4,-18,-214,-900,-10504,-44118,-514714,-24242424,-25221004,-105928218,-1235829214};
static const int hiLim[12]={0,2,30,128,1500,6302,73530,308828,3603000,15132602,176547030,741497528};
static const unsigned char fwdTable[6][7]=
{ /* flowsnake.cpp:46-54: high nibble = next orientation, low = output digit */
  {0x52,0x05,0x06,0x24,0x33,0x40,0x01},
  {0x31,0x10,0x12,0x05,0x43,0x54,0x16},
  {0x46,0x24,0x21,0x10,0x53,0x35,0x22},
  {0x31,0x10,0x03,0x54,0x36,0x35,0x22},
  {0x46,0x24,0x13,0x35,0x42,0x40,0x01},
  {0x52,0x05,0x23,0x40,0x51,0x54,0x16}
};
static const int root1x[6]={1,1,0,-1,-1,0},root1y[6]={0,1,1,0,-1,-1};  /* eisenstein.cpp:51 */

void wbo_to_flowsnake(int n,int *ex,int *ey)
/* toFlowsnake = baseFlow(iToFlowsnake(n)): flowsnake.cpp:92-136.  The transducer maps the
 * 11 base-7 digits of n+1235829214 (most significant first) to the digits of the centred
 * base-7 number m+988663371; m's balanced digits d in [-3,3] then weight powers of 2-w. */
{
  int dig[11],i,ori=0;
  long long v=(long long)n+1235829214LL;
  int x=0,y=0,px=1,py=0;            /* (px,py) = (2-w)^i */
  for (i=0;i<11;i++)
  {
    dig[i]=(int)(v%7);
    v/=7;
  }
  for (i=10;i>=0;i--)
  {
    int t=fwdTable[ori][dig[i]];
    ori=t>>4;
    dig[i]=t&7;
  }
  /* digits dig[i] in [0,6] are balanced digits +3 */
  for (i=0;i<11;i++)
  {
    int d=dig[i]-3,dy=(d+4)/3-1,dx=d-2*dy,nx,ny;
    /* (dx+dy w)*(px+py w), w^2=-1-w : (a,b)*(c,d) = (ac-bd, ad+bc-bd) */
    x+=dx*px-dy*py;
    y+=dx*py+dy*px-dy*py;
    nx=px*2-py*(-1);                 /* (px,py)*(2,-1) */
    ny=px*(-1)+py*2-py*(-1);
    px=nx; py=ny;
  }
  *ex=x;
  *ey=y;
}

int wbo_base_seven(int ex,int ey)
/* baseSeven(): flowsnake.cpp:76-90.  The remainder of e mod (2-w) with least norm is the
 * unique digit d in [-3,3] with d = x+2y (mod 7), since w = 2 (mod 2-w). */
{
  long long ret=0,pow7=1;
  int guard=0;
  while ((ex || ey) && guard++<40)
  {
    int d=(((ex+2*ey)%7)+10)%7-3;
    int dy=(d+4)/3-1,dx=d-2*dy;
    int a=ex-dx,b=ey-dy;
    /* (a+bw)/(2-w) = (a+bw)(3+w)/7 = ((3a-b) + (a+2b)w)/7 */
    ex=(3*a-b)/7;
    ey=(a+2*b)/7;
    ret+=d*pow7;
    pow7*=7;
  }
  return (int)ret;
}

int wbo_from_flowsnake(int ex,int ey,int64_t *n)
/* Inverse of wbo_to_flowsnake (ours; the reference only has the forward map).
 * Returns 0 and the sequence number, or -1 if e needs more than 11 digits. */
{
  int dig[11],i,ori=0,k=0;
  long long v=0;
  for (i=0;i<11;i++)
    dig[i]=3;
  while (ex || ey)
  {
    int d,dy,dx,a,b;
    if (k>=11)
      return -1;
    d=(((ex+2*ey)%7)+10)%7-3;
    dy=(d+4)/3-1; dx=d-2*dy;
    a=ex-dx; b=ey-dy;
    ex=(3*a-b)/7;
    ey=(a+2*b)/7;
    dig[k++]=d+3;
  }
  for (i=10;i>=0;i--)
  {
    int d;
    for (d=0;d<7;d++)
      if ((fwdTable[ori][d]&7)==dig[i])
        break;
    if (d==7)
      return -1;
    ori=fwdTable[ori][d]>>4;
    v=v*7+d;
  }
  *n=v-1235829214LL;
  return 0;
}

int wbo_snake_set_size(double cube_side,double tile_size,double *spacing,int *lo,int *hi)
/* Flowsnake::setSize: flowsnake.cpp:208-230 */
{
  int i,best=0;
  double bestDiff=INFINITY;
  for (i=0;i<12;i++)
  {
    double sp=cube_side/squareSides[i];
    double diff=fabs(log(sp/tile_size));
    if (diff<bestDiff)
    {
      bestDiff=diff;
      best=i;
    }
  }
  *spacing=cube_side/squareSides[best];
  *lo=loLim[best];
  *hi=hiLim[best];
  return best;
}

static void tile_center(int ex,int ey,double spacing,double ccx,double ccy,double *x,double *y)
/* Flowsnake::cyl: flowsnake.cpp:263-271 with Eisenstein -> complex (eisenstein.h:91-94) */
{
  double re=ex-ey/2.,im=ey*SQRT_3_4;
  re*=spacing;
  im*=spacing;
  *x=re+ccx;
  *y=im+ccy;
}

/* ======================= manysum.cpp / matrix.cpp / leastsquares.cpp ====== */

static double tree_sum(const double *a,unsigned len)
{
  if (len==1)
    return a[0];
  return tree_sum(a,len/2)+tree_sum(a+len/2,len/2);
}

double wbo_pairwise_sum(const double *a,unsigned n)
/* pairwisesum(): manysum.cpp:120-154.  The binary-counter merge there amounts to: split
 * a[0..n) into aligned power-of-two blocks by the set bits of n, largest block first;
 * each block is summed as a perfect binary tree; block totals are accumulated from the
 * smallest block (at the end of the array) to the largest, starting from 0. */
{
  double s=0;
  unsigned bit,pos=n;
  for (bit=1;bit && bit<=n;bit<<=1)
    if (n&bit)
    {
      pos-=bit;
      s+=tree_sum(a+pos,bit);
    }
  return s;
}

#define MAXC 4
typedef struct { int rows,cols; double e[MAXC][MAXC]; } smat;

static void sm_swap(smat *m,int r0,int r1)
{
  double t[MAXC];
  memcpy(t,m->e[r0],sizeof(t));
  memcpy(m->e[r0],m->e[r1],sizeof(t));
  memcpy(m->e[r1],t,sizeof(t));
}

static void sm_rowop(smat *a,smat *b,int row0,int row1,int piv)
/* matrix::rowop: matrix.cpp:262-349 (swap / normalise row0 / eliminate from row1) */
{
  int i,flags=0,pivot;
  double slope=0,minslope=INFINITY,detfactor;
  double *rw0=a->e[row0],*rw1=a->e[row1];
  if (piv>=0 && rw0[piv]==0 && rw1[piv]==0)
    piv=-1;
  pivot=piv;
  if (piv>=0 && rw0[piv]==0)
    flags=9;
  for (i=0;piv<0 && i<a->cols;i++)
    if (rw0[i]!=0 || rw1[i]!=0)
    {
      if (fabs(rw0[i])>fabs(rw1[i]) || row0>=row1)
      {
        slope=fabs(rw1[i]/rw0[i]);
        flags&=~8;
      }
      else
      {
        slope=fabs(rw0[i]/rw1[i]);
        flags|=8;
      }
      if (slope<minslope)
      {
        minslope=slope;
        flags=(flags>>3)*9;
        pivot=i;
      }
    }
  flags&=1;
  if (flags)
  {
    sm_swap(a,row0,row1);
    sm_swap(b,row0,row1);
  }
  detfactor=pivot<0?0:rw0[pivot];
  if (detfactor!=0 && detfactor!=1)
  {
    for (i=0;i<a->cols;i++)
      rw0[i]/=detfactor;
    for (i=0;i<b->cols;i++)
      b->e[row0][i]/=detfactor;
  }
  if (pivot>=0)
    slope=rw1[pivot];
  if (slope!=0 && row0!=row1)
  {
    for (i=0;i<a->cols;i++)
      rw1[i]-=rw0[i]*slope;
    for (i=0;i<b->cols;i++)
      b->e[row1][i]-=b->e[row0][i]*slope;
  }
}

static void sm_findpivot(smat *a,smat *b,int row,int column)
/* matrix::findpivot: matrix.cpp:382-422 */
{
  int i,j,pivotrow=-1;
  double maxratio=0,squares[MAXC+1],ratio;
  for (;pivotrow<row && column<a->cols;column++)
    for (i=row;i<a->rows;i++)
    {
      memset(squares,0,sizeof(squares));
      for (j=column+1;j<a->cols;j++)
        squares[j-column-1]=a->e[i][j]*a->e[i][j];
      ratio=(a->e[i][column]*a->e[i][column])/wbo_pairwise_sum(squares,a->cols-column);
      if (ratio>maxratio)
      {
        pivotrow=i;
        maxratio=ratio;
      }
    }
  if (pivotrow>row)
  {
    sm_swap(a,pivotrow,row);
    sm_swap(b,pivotrow,row);
  }
}

static void sm_gausselim(smat *a,smat *b)
/* matrix::gausselim: matrix.cpp:358-380 */
{
  int i,j;
  for (i=0;i<a->rows;i++)
  {
    sm_findpivot(a,b,i,i);
    for (j=0;j<a->rows;j++)
      sm_rowop(a,b,i,j,i);
  }
  for (i=a->rows-1;i>=0;i--)
    for (j=0;j<i;j++)
      sm_rowop(a,b,i,j,i);
}

int wbo_least_squares(const double *a,const double *b,int rows,int cols,double *x)
/* linearLeastSquares: leastsquares.cpp:28-43.  a is rows x cols row-major.  Normal
 * equations with pairwise-summed products (matrix::transmult matrix.cpp:200-215,
 * operator* 181-198), then Gauss-Jordan; zero diagonal -> NaN. */
{
  smat mtm,mtv;
  double *prod;
  int i,j,k;
  if (cols>MAXC || rows<1)
    return -1;
  prod=(double *)malloc(sizeof(double)*rows);
  memset(&mtm,0,sizeof(mtm));
  memset(&mtv,0,sizeof(mtv));
  mtm.rows=mtm.cols=cols;
  mtv.rows=cols;
  mtv.cols=1;
  for (i=0;i<cols;i++)
    for (j=0;j<=i;j++)
    {
      for (k=0;k<rows;k++)
        prod[k]=a[k*cols+i]*a[k*cols+j];
      mtm.e[i][j]=mtm.e[j][i]=wbo_pairwise_sum(prod,rows);
    }
  for (i=0;i<cols;i++)
  {
    for (k=0;k<rows;k++)
      prod[k]=a[k*cols+i]*b[k];
    mtv.e[i][0]=wbo_pairwise_sum(prod,rows);
  }
  free(prod);
  sm_gausselim(&mtm,&mtv);
  for (i=0;i<cols;i++)
  {
    if (mtm.e[i][i]==0)
      mtv.e[i][0]=NAN;
    x[i]=mtv.e[i][0];
  }
  return 0;
}

/* ======================= classify.cpp: surround =========================== */

static int cmp_i32(const void *a,const void *b)
{
  int32_t x=*(const int32_t *)a,y=*(const int32_t *)b;
  return (x>y)-(x<y);
}

int wbo_surround(const int32_t *dirs,int n)
/* surround(): classify.cpp:67-94 evaluated on the full direction set: true iff there are
 * at least two distinct directions and no circular gap (mod 2^31) reaches 144 degrees.
 * (The thinning side effect there cannot change the result: SURVEY.md §8a row C3.) */
{
  int32_t *s;
  int i,m=0,ret;
  if (n<2)
    return 0;
  s=(int32_t *)malloc(sizeof(int32_t)*n);
  memcpy(s,dirs,sizeof(int32_t)*n);
  qsort(s,n,sizeof(int32_t),cmp_i32);
  for (i=0;i<n;i++)
    if (i==0 || s[i]!=s[i-1])
      s[m++]=s[i];
  ret=m>1;
  for (i=1;i<m;i++)
    if ((((uint32_t)s[i]-(uint32_t)s[i-1])&INT_MAX)>=DEG144)
      ret=0;
  if ((((uint32_t)s[0]-(uint32_t)s[m-1])&INT_MAX)>=DEG144)
    ret=0;
  free(s);
  return ret;
}

/* ======================= ldecimal.cpp ===================================== */

int wbo_ldecimal(double x,char *out,int outlen)
/* ldecimal(x,0): ldecimal.cpp:31-127 — fewest significant digits that read back equal,
 * then a plain/exponent layout chosen by the exponent. */
{
  char buf[64],mant[64],res[96];
  int prec,iexp,i,n,neg=0;
  char *e;
  for (prec=0;prec<=DBL_DIG+3;prec++)
  {
    snprintf(buf,sizeof(buf),"%.*e",prec,x);
    if (atof(buf)==x)
      break;
  }
  e=strchr(buf,'e');
  iexp=atoi(e+1);
  *e=0;
  /* digits of the significand without sign and dot, trailing zeros removed */
  n=0;
  for (i=0;buf[i];i++)
    if (buf[i]=='-')
      neg=1;
    else if (buf[i]!='.')
      mant[n++]=buf[i];
  while (n>1 && mant[n-1]=='0')
    n--;
  mant[n]=0;
  /* mant = d0 d1 d2...; value = d0.d1d2... * 10^iexp */
  {
    char m[64]="",a[64]="";
    int ml,al;
    m[0]=mant[0]; m[1]=0;
    strcpy(a,mant+1);
    if (n==1 && mant[0]=='0')
      a[0]=0;
    ml=1; al=(int)strlen(a);
    if (iexp<0 && iexp>-5)
    {
      memmove(a+1,a,al+1);
      a[0]=m[0];
      al++;
      m[0]=0;
      ml=0;
      iexp++;
    }
    if (iexp>0)
    {
      int ch=iexp>al?al:iexp;
      strncat(m,a,ch);
      ml+=ch;
      memmove(a,a+ch,al-ch+1);
      al-=ch;
      iexp-=ch;
    }
    while (iexp>-5 && iexp<0 && ml==0)
    {
      memmove(a+1,a,al+1);
      a[0]='0';
      al++;
      iexp++;
    }
    while (iexp<3 && iexp>0 && al==0)
    {
      m[ml++]='0';
      m[ml]=0;
      iexp--;
    }
    /* a keeps trailing zeros only if they were moved in front; strip as the reference does
     * before layout (find_last_not_of('0') happens earlier there, so nothing to do here) */
    res[0]=0;
    if (neg)
      strcat(res,"-");
    strcat(res,m);
    if (al)
    {
      strcat(res,".");
      strcat(res,a);
    }
    if (iexp)
      snprintf(res+strlen(res),16,"e%d",iexp);
  }
  if ((int)strlen(res)+1>outlen)
    return -1;
  strcpy(out,res);
  return (int)strlen(res);
}

/* ======================= las.cpp: record decode =========================== */

int wbo_decode(const uint8_t *recs,uint64_t n,int fmt,int rec_len,int32_t *xyz,uint8_t *cls,uint8_t *ret_num)
/* LasHeader::readPoint: las.cpp:735-775 — X,Y,Z little-endian int32 at 0,4,8; formats 0-5:
 * return number = byte14&7, class = byte15&31; formats 6-10: return = byte14&15, class = byte16. */
{
  uint64_t i;
  if (fmt<0 || fmt>10 || rec_len<20)
    return -1;
  for (i=0;i<n;i++)
  {
    const uint8_t *r=recs+i*(uint64_t)rec_len;
    int k;
    for (k=0;k<3;k++)
      xyz[3*i+k]=(int32_t)((uint32_t)r[4*k]|((uint32_t)r[4*k+1]<<8)|((uint32_t)r[4*k+2]<<16)|((uint32_t)r[4*k+3]<<24));
    if (fmt<6)
    {
      ret_num[i]=r[14]&7;
      cls[i]=r[15]&31;
    }
    else
    {
      ret_num[i]=r[14]&15;
      cls[i]=r[16];
    }
  }
  return 0;
}

void wbo_coords(const int32_t *xyz,uint64_t n,const double scale[3],const double offset[3],double unit,double *out)
/* las.cpp:808: location = (offset + scale*int) * unit, un-fused */
{
  uint64_t i;
  int k;
  for (i=0;i<n;i++)
    for (k=0;k<3;k++)
    {
      double p=scale[k]*xyz[3*i+k];
      out[3*i+k]=(offset[k]+p)*unit;
    }
}

/* ======================= octree.cpp ======================================= */

void wbo_size_fit(const double *c,int n,double center[3],double *side_out)
/* Octree::sizeFit: octree.cpp:268-310 */
{
  double mn[3]={HUGE_VAL,HUGE_VAL,HUGE_VAL},mx[3]={-HUGE_VAL,-HUGE_VAL,-HUGE_VAL};
  double side,x,y,z;
  int i,k;
  for (i=0;i<n;i++)
    for (k=0;k<3;k++)
    {
      if (c[3*i+k]>mx[k]) mx[k]=c[3*i+k];
      if (c[3*i+k]<mn[k]) mn[k]=c[3*i+k];
    }
  if (mx[2]<=mn[2] && mx[1]<=mn[1] && mx[0]<=mn[0])
  {
    *side_out=0;
    center[0]=center[1]=center[2]=0;
    return;
  }
  side=(mx[0]+mx[1]+mx[2]-mn[0]-mn[1]-mn[2])/3;
  side/=significand(side);
  x=mn[0]-side;
  y=mn[1]-side;
  z=mn[2]-side;
  while (x+side<=mx[0] || y+side<=mx[1] || z+side<=mx[2])
  {
    side*=2;
    x=(rint((mn[0]+mx[0])/side*8)-8)*side/16;
    y=(rint((mn[1]+mx[1])/side*8)-8)*side/16;
    z=(rint((mn[2]+mx[2])/side*8)-8)*side/16;
  }
  center[0]=x+side/2;
  center[1]=y+side/2;
  center[2]=z+side/2;
  *side_out=side;
}

void wbo_bbox_cube(const double *c,int n,double cube[4])
/* The cube handed to Flowsnake::setSize: wolkencanvas.cpp:502-519 with BoundRect::include
 * (boundrect.cpp:60-73) at orientation 0, where xy::dirbound (point.cpp:85-93) uses the
 * long-double-derived sin/cos of 0, 90, 180, 270 degrees (not exactly 0 and 1). */
{
  double b[6]={INFINITY,INFINITY,INFINITY,INFINITY,INFINITY,INFINITY};
  double left,bottom,right,top,low,high,side;
  int i,k;
  for (i=0;i<n;i++)
  {
    for (k=0;k<4;k++)
    {
      int ang=(int)((unsigned)k*DEG90);
      double s=(double)sinl(ang*M_PIl/1073741824.),co=(double)cosl(ang*M_PIl/1073741824.);
      double v=c[3*i]*co+c[3*i+1]*s;
      if (v<b[k]) b[k]=v;
    }
    if (c[3*i+2]<b[4]) b[4]=c[3*i+2];
    if (-c[3*i+2]<b[5]) b[5]=-c[3*i+2];
  }
  left=b[0]; bottom=b[1]; right=-b[2]; top=-b[3]; low=b[4]; high=-b[5];
  side=right-left;
  if (top-bottom>side) side=top-bottom;
  if (high-low>side) side=high-low;
  cube[0]=(right+left)/2;
  cube[1]=(top+bottom)/2;
  cube[2]=(high+low)/2;
  cube[3]=side;
}

uint64_t wbo_morton_key(const double p[3],const double center[3],double side)
/* 21 steps of Octree::findBlock's descent (octree.cpp:199-216): child = z*4+y*2+x with
 * bit = coordinate >= centre; child centre = centre +- side/4 (octree.cpp:335-337). */
{
  uint64_t key=0;
  double cx=center[0],cy=center[1],cz=center[2],q=side/4;
  int l;
  for (l=0;l<WBO_LEVELS;l++)
  {
    int xb=p[0]>=cx,yb=p[1]>=cy,zb=p[2]>=cz;
    key=(key<<3)|(uint64_t)(zb*4+yb*2+xb);
    cx+=(2*xb-1)*q;
    cy+=(2*yb-1)*q;
    cz+=(2*zb-1)*q;
    q/=2;
  }
  return key;
}

typedef struct { uint64_t key; uint32_t idx; } keyidx;

static int cmp_keyidx(const void *a,const void *b)
{
  const keyidx *x=(const keyidx *)a,*y=(const keyidx *)b;
  if (x->key!=y->key)
    return x->key<y->key?-1:1;
  return (x->idx>y->idx)-(x->idx<y->idx);
}

int wbo_sort(const double *pts,uint64_t n,const double center[3],double side,uint64_t *keys_sorted,uint32_t *order)
{
  keyidx *ki=(keyidx *)malloc(sizeof(keyidx)*(n?n:1));
  uint64_t i;
  if (!ki)
    return -1;
  #pragma omp parallel for
  for (i=0;i<n;i++)
  {
    ki[i].key=wbo_morton_key(pts+3*i,center,side);
    ki[i].idx=(uint32_t)i;
  }
  qsort(ki,n,sizeof(keyidx),cmp_keyidx);
  for (i=0;i<n;i++)
  {
    keys_sorted[i]=ki[i].key;
    order[i]=ki[i].idx;
  }
  free(ki);
  return 0;
}

static void split_rec(const uint64_t *keys,uint64_t lo,uint64_t hi,int depth,uint64_t prefix,
                      const double c[3],double side,wbo_leaf *out,int64_t cap,int64_t *cnt)
/* A cube that ever held more than 537 (distinct) points is an internal node
 * (OctStore::put -> split, octree.cpp:849-876, 1295-1338; Octree::split 312-346); its
 * non-empty children appear in child-index order (Octree::dump, octree.cpp:360-376). */
{
  int ch;
  uint64_t pos=lo;
  (void)prefix;
  for (ch=0;ch<8;ch++)
  {
    int shift=3*(WBO_LEVELS-1-depth);
    uint64_t end=pos;
    double cc[3],h=side/4;
    while (end<hi && ((keys[end]>>shift)&7)==(uint64_t)ch)
      end++;
    if (end==pos)
      continue;
    cc[0]=c[0]+((ch&1)?h:-h);
    cc[1]=c[1]+((ch&2)?h:-h);
    cc[2]=c[2]+((ch&4)?h:-h);
    if (end-pos<=WBO_RECORDS || depth+1>=WBO_LEVELS)
    {
      if (*cnt<cap)
      {
        wbo_leaf *l=out+*cnt;
        l->first=pos;
        l->count=(uint32_t)(end-pos);
        l->depth=depth+1;
        l->cx=cc[0]; l->cy=cc[1]; l->cz=cc[2];
        l->half=side/4;
      }
      (*cnt)++;
    }
    else
      split_rec(keys,pos,end,depth+1,0,cc,side/2,out,cap,cnt);
    pos=end;
  }
}

int64_t wbo_leaves(const uint64_t *keys_sorted,uint64_t n,const double center[3],double side,wbo_leaf *out,int64_t cap)
{
  int64_t cnt=0;
  if (n)
    split_rec(keys_sorted,0,n,0,0,center,side,out,cap,&cnt);
  return cnt;
}

int64_t wbo_dump(const wbo_leaf *leaves,int64_t n_leaves,char *buf,int64_t buflen)
/* OctBuffer::dump / OctStore::dump: octree.cpp:673-689, 888-891 */
{
  int64_t i,pos=0;
  uint64_t total=0;
  char a[40],b[40],c[40],d[40],line[256];
  for (i=0;i<n_leaves;i++)
  {
    int len;
    wbo_ldecimal(leaves[i].cx,a,40);
    wbo_ldecimal(leaves[i].cy,b,40);
    wbo_ldecimal(leaves[i].cz,c,40);
    wbo_ldecimal(leaves[i].half,d,40);
    len=snprintf(line,sizeof(line),"(%s,%s,%s)\xc2\xb1%s %u points\n",a,b,c,d,leaves[i].count);
    if (pos+len>=buflen)
      return -1;
    memcpy(buf+pos,line,len);
    pos+=len;
    total+=leaves[i].count;
  }
  {
    int len=snprintf(line,sizeof(line),"%llu total points\n",(unsigned long long)total);
    if (pos+len>=buflen)
      return -1;
    memcpy(buf+pos,line,len);
    pos+=len;
  }
  buf[pos]=0;
  return pos;
}

/* ======================= tiles: membership ================================ */

typedef struct
{
  double spacing,ccx,ccy,radius;
  int lo,hi;
} snake_t;

static void snake_init(snake_t *s,const double cube[4],double tile_size)
{
  wbo_snake_set_size(cube[3],tile_size,&s->spacing,&s->lo,&s->hi);
  s->ccx=cube[0];
  s->ccy=cube[1];
  s->radius=s->spacing*41/71;       /* flowsnake.cpp:265 */
}

static int covering_tiles(const snake_t *s,double px,double py,int64_t *ns,int *exs,int *eys)
/* All tiles whose cylinder contains (px,py) (Cylinder::in, shape.cpp:214-218) and whose
 * sequence number lies in the snake's range.  Candidates: the lattice point obtained by
 * rounding plus its two surrounding rings (19 addresses) — a superset of every centre within
 * 41/71 spacing, however the rounding falls.  Result is NOT ordered. */
{
  double u=(px-s->ccx)/s->spacing,v=(py-s->ccy)/s->spacing;
  int y0=(int)lrint(v/SQRT_3_4),x0=(int)lrint(u+y0*0.5),dx,dy,cnt=0;
  for (dy=-2;dy<=2;dy++)
    for (dx=-2;dx<=2;dx++)
    {
      int ex=x0+dx,ey=y0+dy;
      double cx,cy;
      int64_t n;
      if (dx-dy>2 || dy-dx>2)        /* keep the 19 addresses of hex-norm <= 2 */
        continue;
      tile_center(ex,ey,s->spacing,s->ccx,s->ccy,&cx,&cy);
      if (!(hypot(cx-px,cy-py)<=s->radius))
        continue;
      if (wbo_from_flowsnake(ex,ey,&n) || n<s->lo || n>s->hi)
        continue;
      ns[cnt]=n;
      exs[cnt]=ex;
      eys[cnt]=ey;
      cnt++;
    }
  return cnt;
}

/* ======================= scan.cpp ========================================= */

typedef struct { int64_t n; uint64_t k; int ex,ey; } member;

static int cmp_member(const void *a,const void *b)
{
  const member *x=(const member *)a,*y=(const member *)b;
  if (x->n!=y->n)
    return x->n<y->n?-1:1;
  return (x->k>y->k)-(x->k<y->k);
}

static void scan_tile(const double *pts,const member *m,uint64_t cnt,const snake_t *s,
                      double min_hyp,wbo_tile *t)
/* scanCylinder: scan.cpp:31-140 for one tile; m[0..cnt) are its points in canonical order */
{
  double ccx,ccy,*a,*b,*zu,sl[3],sx,sy,len;
  double bottom=INFINITY,bottom2=INFINITY,top=-INFINITY,density=0;
  uint64_t i,nBottom=0;
  int histo[7]={0,0,0,0,0,0,0},treeFlags=0,j;
  tile_center(m[0].ex,m[0].ey,s->spacing,s->ccx,s->ccy,&ccx,&ccy);
  a=(double *)malloc(sizeof(double)*3*cnt);
  b=(double *)malloc(sizeof(double)*cnt);
  zu=(double *)malloc(sizeof(double)*cnt);
  for (i=0;i<cnt;i++)
  {
    const double *p=pts+3*m[i].k;
    a[3*i]=p[0]-ccx;
    a[3*i+1]=p[1]-ccy;
    a[3*i+2]=1;
    b[i]=p[2]-0;
  }
  wbo_least_squares(a,b,(int)cnt,3,sl);
  sx=sl[0];
  sy=sl[1];
  len=hypot(sx,sy);
  if (len>1)
  {
    double l2=hypot(sx,sy);
    sx/=l2;
    sy/=l2;
  }
  if (isnan(sx) || isnan(sy))
    sx=sy=0;
  for (i=0;i<cnt;i++)
  {
    double z=sy*a[3*i+1]+sx*a[3*i];     /* dot(): a.y*b.y+a.x*b.x, point.cpp:199-202 */
    zu[i]=b[i]-z;
    if (zu[i]<bottom)
    {
      bottom2=bottom;
      bottom=zu[i];
    }
    if (zu[i]>top)
      top=zu[i];
  }
  if (isinf(bottom2))
    bottom2=bottom;
  for (i=0;i<cnt;i++)
    if (zu[i]<bottom2+2*s->radius)
    {
      double x=a[3*i],y=a[3*i+1];
      int sector=(int)lrint(atan2(y,x)*3/M_PI);
      if (sector<0)
        sector+=6;
      sector=(sector%6)+1;
      if (hypot(x,y)<s->radius/SQRT7)
        sector=0;
      histo[sector]++;
      nBottom++;
    }
  for (j=0;j<7;j++)
    density+=histo[j]*histo[j];
  if (cnt>nBottom && density<7)
    treeFlags=1;
  density=sqrt(density)*SQRT7/(s->radius*s->radius)/M_PI;
  if (cnt>nBottom && density<0.5)
    treeFlags=1;
  if (top-bottom>1.5)
    treeFlags=1;
  t->n=(int32_t)m[0].n;
  t->ex=m[0].ex;
  t->ey=m[0].ey;
  t->nPoints=(int32_t)cnt;
  t->treeFlags=treeFlags;
  t->density=density;
  t->hyperboloidSize=sqrt(1/density+min_hyp*min_hyp);
  t->height=top-bottom;
  free(a);
  free(b);
  free(zu);
}

int64_t wbo_scan(const double *pts,uint64_t n,const double cube[4],double tile_size,
                 double min_hyp,wbo_tile *out,int64_t cap)
{
  snake_t s;
  member *mem;
  uint64_t i,nm=0,capm=n*4+16,start;
  int64_t nt=0;
  snake_init(&s,cube,tile_size);
  mem=(member *)malloc(sizeof(member)*capm);
  for (i=0;i<n;i++)
  {
    int64_t ns[19];
    int exs[19],eys[19],c=covering_tiles(&s,pts[3*i],pts[3*i+1],ns,exs,eys),j;
    for (j=0;j<c;j++)
    {
      if (nm==capm)
      {
        capm*=2;
        mem=(member *)realloc(mem,sizeof(member)*capm);
      }
      mem[nm].n=ns[j];
      mem[nm].k=i;
      mem[nm].ex=exs[j];
      mem[nm].ey=eys[j];
      nm++;
    }
  }
  qsort(mem,nm,sizeof(member),cmp_member);
  for (start=0;start<nm;)
  {
    uint64_t end=start;
    while (end<nm && mem[end].n==mem[start].n)
      end++;
    if (nt<cap)
      scan_tile(pts,mem+start,end-start,&s,min_hyp,out+nt);
    nt++;
    start=end;
  }
  free(mem);
  return nt;
}

static int cmp_tile_addr(const void *a,const void *b)
{
  const wbo_tile *x=(const wbo_tile *)a,*y=(const wbo_tile *)b;
  if (x->ey!=y->ey)
    return x->ey<y->ey?-1:1;
  return (x->ex>y->ex)-(x->ex<y->ex);
}

static const wbo_tile *find_tile(const wbo_tile *byaddr,int64_t n,int ex,int ey)
{
  wbo_tile key;
  key.ex=ex;
  key.ey=ey;
  return (const wbo_tile *)bsearch(&key,byaddr,n,sizeof(wbo_tile),cmp_tile_addr);
}

int wbo_postscan(wbo_tile *tiles,int64_t n_tiles,double spacing)
/* postscanCylinder: scan.cpp:142-179.  A tile absent from the table has nPoints==0.
 * Only treeFlags/nPoints of OTHER tiles are read, so the visiting order is immaterial. */
{
  wbo_tile *byaddr=(wbo_tile *)malloc(sizeof(wbo_tile)*(n_tiles?n_tiles:1));
  int64_t t;
  memcpy(byaddr,tiles,sizeof(wbo_tile)*n_tiles);
  qsort(byaddr,n_tiles,sizeof(wbo_tile),cmp_tile_addr);
  for (t=0;t<n_tiles;t++)
  {
    wbo_tile *th=tiles+t;
    int i=1,j,nontree,ringcount,count=0;
    if (!th->nPoints)
      continue;
    do
    {
      for (nontree=ringcount=j=0;j<6 && (th->treeFlags&1);j++)
      {
        const wbo_tile *o=find_tile(byaddr,n_tiles,th->ex+root1x[j]*i,th->ey+root1y[j]*i);
        if (o && o->nPoints)
        {
          ringcount++;
          if (o->treeFlags&1)
            count++;
          else
            nontree++;
        }
      }
      ++i;
    } while (ringcount && !nontree);
    {
      double c=count*spacing/6;
      th->hyperboloidSize=sqrt(th->hyperboloidSize*th->hyperboloidSize+c*c);
    }
  }
  free(byaddr);
  return 0;
}

/* ======================= classify.cpp ===================================== */

static int cmp_tile_n(const void *a,const void *b)
{
  const wbo_tile *x=(const wbo_tile *)a,*y=(const wbo_tile *)b;
  return (x->n>y->n)-(x->n<y->n);
}

int wbo_point_hyperboloid_sizes(const double *pts,uint64_t n,const double cube[4],double tile_size,
                                const wbo_tile *tiles,int64_t n_tiles,double *out)
/* The hyperboloidSize classifyCylinder uses for each point (classify.cpp:121-131): that of the LAST tile in
 * flowsnake order whose cylinder contains it; NaN for a point in no tile.  Lets a test drive the classify
 * step alone. */
{
  snake_t s;
  wbo_tile *byn;
  uint64_t i;
  snake_init(&s,cube,tile_size);
  byn=(wbo_tile *)malloc(sizeof(wbo_tile)*(n_tiles?n_tiles:1));
  memcpy(byn,tiles,sizeof(wbo_tile)*n_tiles);
  qsort(byn,n_tiles,sizeof(wbo_tile),cmp_tile_n);
  for (i=0;i<n;i++)
  {
    int64_t ns[19],best=LLONG_MIN;
    int exs[19],eys[19],c=covering_tiles(&s,pts[3*i],pts[3*i+1],ns,exs,eys),j;
    const wbo_tile *t=NULL;
    for (j=0;j<c;j++)
      if (ns[j]>best)
        best=ns[j];
    if (c)
    {
      wbo_tile key;
      key.n=(int32_t)best;
      t=(const wbo_tile *)bsearch(&key,byn,n_tiles,sizeof(wbo_tile),cmp_tile_n);
    }
    out[i]=t?t->hyperboloidSize:NAN;
  }
  free(byn);
  return 0;
}

static int classify_impl(const double *pts,uint64_t n,const double cube[4],double tile_size,
                 double max_slope,double thickness,const wbo_tile *tiles,int64_t n_tiles,
                 const uint64_t *sel,uint64_t n_sel,uint8_t *labels,uint64_t *margin_count)
/* classifyCylinder: classify.cpp:96-173, as a pure per-point function:
 *   tile  = the LAST tile in flowsnake order whose cylinder contains P (1-thread reference:
 *           a later tile re-classifies and overwrites, classify.cpp:158-165);
 *   H     = Hyperboloid(P-(0,0,thickness), tile.hyperboloidSize, maxSlope);
 *   label = 1 if the bearings dir(P,Q) of all Q in H with dist_xy(P,Q)!=0 surround P, else 2.
 * The candidate search uses a uniform xy grid with per-cell minimum z; a cell is skipped only
 * if even its lowest point at its nearest xy could not be inside H (with slack), so every
 * point of the cloud that can be in H is tested with the reference's exact predicate. */
{
  snake_t s;
  wbo_tile *byn;
  double minx=INFINITY,miny=INFINITY,maxx=-INFINITY,maxy=-INFINITY,minz=INFINITY,cell;
  int64_t gx,gy,*cellStart,ncell;
  uint32_t *cellPts;
  double *cellMinZ;
  uint64_t i,margins=0,si;
  const uint64_t n_query=sel?n_sel:n;   /* sel: classify only these points (each still against the WHOLE cloud) */
  snake_init(&s,cube,tile_size);
  if (!tablesFilled)
    wbo_fill_tan_tables();
  byn=(wbo_tile *)malloc(sizeof(wbo_tile)*(n_tiles?n_tiles:1));
  memcpy(byn,tiles,sizeof(wbo_tile)*n_tiles);
  qsort(byn,n_tiles,sizeof(wbo_tile),cmp_tile_n);
  for (i=0;i<n;i++)
  {
    if (pts[3*i]<minx) minx=pts[3*i];
    if (pts[3*i]>maxx) maxx=pts[3*i];
    if (pts[3*i+1]<miny) miny=pts[3*i+1];
    if (pts[3*i+1]>maxy) maxy=pts[3*i+1];
    if (pts[3*i+2]<minz) minz=pts[3*i+2];
  }
  cell=sqrt((maxx-minx+1e-9)*(maxy-miny+1e-9)/((double)n/16+1));
  if (!(cell>0))
    cell=1;
  gx=(int64_t)((maxx-minx)/cell)+1;
  gy=(int64_t)((maxy-miny)/cell)+1;
  ncell=gx*gy;
  cellStart=(int64_t *)calloc(ncell+1,sizeof(int64_t));
  cellMinZ=(double *)malloc(sizeof(double)*ncell);
  cellPts=(uint32_t *)malloc(sizeof(uint32_t)*(n?n:1));
  for (i=0;i<(uint64_t)ncell;i++)
    cellMinZ[i]=INFINITY;
  #define CELL_OF(px,py) (((int64_t)(((py)-miny)/cell))*gx+(int64_t)(((px)-minx)/cell))
  for (i=0;i<n;i++)
    cellStart[CELL_OF(pts[3*i],pts[3*i+1])+1]++;
  for (i=0;i<(uint64_t)ncell;i++)
    cellStart[i+1]+=cellStart[i];
  {
    int64_t *fill=(int64_t *)malloc(sizeof(int64_t)*ncell);
    memcpy(fill,cellStart,sizeof(int64_t)*ncell);
    for (i=0;i<n;i++)
    {
      int64_t c=CELL_OF(pts[3*i],pts[3*i+1]);
      cellPts[fill[c]++]=(uint32_t)i;
      if (pts[3*i+2]<cellMinZ[c])
        cellMinZ[c]=pts[3*i+2];
    }
    free(fill);
  }
  #pragma omp parallel reduction(+:margins)
  {
    int32_t *dirs=NULL;
    int dcap=0;
    #pragma omp for schedule(dynamic,256)
    for (si=0;si<n_query;si++)
    {
      const uint64_t i=sel?sel[si]:si;
      const double *P=pts+3*i;
      int64_t ns[19],best=LLONG_MIN;
      int exs[19],eys[19],c=covering_tiles(&s,P[0],P[1],ns,exs,eys),j,nd=0,marg=0;
      const wbo_tile *t=NULL;
      double r,por,por2,vz,cz,reach;
      int64_t cx0,cy0,ix,iy,rad;
      for (j=0;j<c;j++)
        if (ns[j]>best)
          best=ns[j];
      if (c)
      {
        wbo_tile key;
        key.n=(int32_t)best;
        t=(const wbo_tile *)bsearch(&key,byn,n_tiles,sizeof(wbo_tile),cmp_tile_n);
      }
      if (!t)
      {
        labels[si]=0;     /* in no tile: the reference never classifies it */
        continue;
      }
      r=t->hyperboloidSize;
      por=r*(max_slope*max_slope);
      por2=por*por;
      vz=P[2]-thickness;
      cz=vz+por;
      /* zd^2-(d s)^2>=por2 and zd>0  =>  d <= sqrt(zd^2-por2)/s <= zd/s; zd <= cz-zmin */
      cx0=(int64_t)((P[0]-minx)/cell);
      cy0=(int64_t)((P[1]-miny)/cell);
      for (rad=0;;rad++)
      {
        int any=0;
        for (iy=cy0-rad;iy<=cy0+rad;iy++)
          for (ix=cx0-rad;ix<=cx0+rad;ix++)
          {
            int64_t cc,k;
            double nx,ny,dd,lim;
            if (ix<0 || iy<0 || ix>=gx || iy>=gy)
              continue;
            if (ix!=cx0-rad && ix!=cx0+rad && iy!=cy0-rad && iy!=cy0+rad)
              continue;
            any=1;
            cc=iy*gx+ix;
            if (cellStart[cc]==cellStart[cc+1])
              continue;
            /* nearest xy of the cell rectangle to P */
            nx=P[0]; ny=P[1];
            if (nx<minx+ix*cell) nx=minx+ix*cell; else if (nx>minx+(ix+1)*cell) nx=minx+(ix+1)*cell;
            if (ny<miny+iy*cell) ny=miny+iy*cell; else if (ny>miny+(iy+1)*cell) ny=miny+(iy+1)*cell;
            dd=hypot(nx-P[0],ny-P[1])-1e-6;
            if (dd<0) dd=0;
            lim=cz-cellMinZ[cc];
            if (lim<=0 || lim*lim-(dd*max_slope)*(dd*max_slope)<por2*(1-1e-9)-1e-9)
              continue;
            for (k=cellStart[cc];k<cellStart[cc+1];k++)
            {
              const double *Q=pts+3*cellPts[k];
              double d=hypot(P[0]-Q[0],P[1]-Q[1]);
              double zd=cz-Q[2],ds=d*max_slope,lhs=zd*zd-ds*ds;
              int in=max_slope>0?(zd>0 && lhs>=por2):(zd<0 && lhs>=por2);
              if (fabs(lhs-por2)<=1e-12*(zd*zd+por2) && d!=0)
                marg=1;
              if (in && d!=0)
              {
                if (nd==dcap)
                {
                  dcap=dcap?dcap*2:256;
                  dirs=(int32_t *)realloc(dirs,sizeof(int32_t)*dcap);
                }
                dirs[nd++]=wbo_atan2i(Q[1]-P[1],Q[0]-P[0]);   /* dir(a,b)=atan2i(b-a) */
              }
            }
          }
        if (!any)
          break;
        /* beyond this ring the nearest distance is rad*cell; stop when even the global
         * lowest point could not be inside */
        reach=(double)rad*cell;
        {
          double lim=cz-minz+1e-6;          /* no point is lower than minz */
          if (lim<=0 || lim*lim-(reach*max_slope)*(reach*max_slope)<por2*(1-1e-9)-1e-9)
            break;
        }
      }
      labels[si]=wbo_surround(dirs,nd)?1:2;
      margins+=marg;
    }
    free(dirs);
  }
  if (margin_count)
    *margin_count=margins;
  free(byn);
  free(cellStart);
  free(cellMinZ);
  free(cellPts);
  return 0;
}

int wbo_classify(const double *pts,uint64_t n,const double cube[4],double tile_size,
                 double max_slope,double thickness,const wbo_tile *tiles,int64_t n_tiles,
                 uint8_t *labels,uint64_t *margin_count)
{
  return classify_impl(pts,n,cube,tile_size,max_slope,thickness,tiles,n_tiles,NULL,0,labels,margin_count);
}

int wbo_classify_sel(const double *pts,uint64_t n,const double cube[4],double tile_size,
                     double max_slope,double thickness,const wbo_tile *tiles,int64_t n_tiles,
                     const uint64_t *sel,uint64_t n_sel,uint8_t *labels,uint64_t *margin_count)
/* the same test for the points sel[0..n_sel) only (positions in pts), labels[k] for sel[k]: the whole cloud is
 * still searched for every one of them, so a sample of a cloud too large to classify in full can be checked */
{
  return classify_impl(pts,n,cube,tile_size,max_slope,thickness,tiles,n_tiles,sel,n_sel,labels,margin_count);
}
