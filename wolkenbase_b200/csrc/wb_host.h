// wb_host.h — host-side arithmetic shared by the C ABI and the C++ shims: the pieces of the
// reference that run once per job on the CPU (Octree::sizeFit, the bounding cube handed to
// Flowsnake::setSize, fillTanTables, ldecimal).  Plain C++, no CUDA.  Must be compiled without
// FMA contraction (x86-64 default).
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cfloat>
#include <string>
#include <vector>

namespace wbhost
{

static const double kSquareSides[12]=
{ // flowsnake.cpp:30-44: biggest square inside a Gosper island of order i
  0.6583539906808145,1.8501627472990723,4.286014912881196,12.6716160058597,
  32.85016274729906,79.01914481623778,243.0343734125204,592.8501627472989,
  1510.850162747299,4399.956311577825,10731.850162747294,29198.8501627473
};
static const int kLoLim[12]={0,-4,-18,-214,-900,-10504,-44118,-514714,-24242424,-25221004,-105928218,-1235829214};
static const int kHiLim[12]={0,2,30,128,1500,6302,73530,308828,3603000,15132602,176547030,741497528};
static const unsigned char kFwdTable[6][7]=
{ // flowsnake.cpp:46-54
  {0x52,0x05,0x06,0x24,0x33,0x40,0x01},
  {0x31,0x10,0x12,0x05,0x43,0x54,0x16},
  {0x46,0x24,0x21,0x10,0x53,0x35,0x22},
  {0x31,0x10,0x03,0x54,0x36,0x35,0x22},
  {0x46,0x24,0x13,0x35,0x42,0x40,0x01},
  {0x52,0x05,0x23,0x40,0x51,0x54,0x16}
};

inline void fillFlowsnakeTables(unsigned char out[96])
// [ori*8+d] = kFwdTable (next orientation<<4 | output digit); [48+ori*8+digit] = its inverse per orientation
// (next orientation<<4 | input digit d): every row's output digits are a permutation of 0..6
{
  for (int i=0;i<96;i++)
    out[i]=0;
  for (int i=0;i<6;i++)
    for (int j=0;j<7;j++)
    {
      out[i*8+j]=kFwdTable[i][j];
      out[48+i*8+(kFwdTable[i][j]&7)]=(unsigned char)((kFwdTable[i][j]&0xf0)|j);
    }
}

inline void fillTanTables(double *tanT /*512*/,double *cosT /*512*/,double *sinT /*512*/)
// angle.cpp:305-320 via sin/cos/tan(int) (angle.cpp:45-63): long double libm, rounded to double
{
  for (int i=0;i<511;i++)
    tanT[i]=(double)tanl((i*0x100000-0xff00000)*M_PIl/1073741824.);
  tanT[511]=0;
  for (int i=0;i<512;i++)
  {
    sinT[i]=(double)sinl((i*0x100000-0xff80000)*M_PIl/1073741824.);
    cosT[i]=(double)cosl((i*0x100000-0xff80000)*M_PIl/1073741824.);
  }
}

inline void sizeFit(const double *c,int n,double center[3],double *sideOut)
// Octree::sizeFit, octree.cpp:268-310: power-of-two side, corner on a side/16 grid
{
  double mn[3]={HUGE_VAL,HUGE_VAL,HUGE_VAL},mx[3]={-HUGE_VAL,-HUGE_VAL,-HUGE_VAL};
  for (int i=0;i<n;i++)
    for (int k=0;k<3;k++)
    {
      if (c[3*i+k]>mx[k]) mx[k]=c[3*i+k];
      if (c[3*i+k]<mn[k]) mn[k]=c[3*i+k];
    }
  center[0]=center[1]=center[2]=0;
  *sideOut=0;
  if (mx[2]<=mn[2] && mx[1]<=mn[1] && mx[0]<=mn[0])
    return;
  double side=(mx[0]+mx[1]+mx[2]-mn[0]-mn[1]-mn[2])/3;
  side/=significand(side);
  double x=mn[0]-side,y=mn[1]-side,z=mn[2]-side;
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
  *sideOut=side;
}

inline void boundRect(const double *c,int n,double box[6])
// BoundRect::include at orientation 0 (boundrect.cpp:60-73) over corners: {left,bottom,low,right,top,high}
{
  double b[6];
  for (int k=0;k<6;k++)
    b[k]=INFINITY;
  for (int i=0;i<n;i++)
  {
    for (int k=0;k<4;k++)
    {
      int ang=(int)((unsigned)k*0x20000000u);
      double s=(double)sinl(ang*M_PIl/1073741824.),co=(double)cosl(ang*M_PIl/1073741824.);
      double v=c[3*i]*co+c[3*i+1]*s;
      if (v<b[k]) b[k]=v;
    }
    if (c[3*i+2]<b[4]) b[4]=c[3*i+2];
    if (-c[3*i+2]<b[5]) b[5]=-c[3*i+2];
  }
  box[0]=b[0]; box[1]=b[1]; box[2]=b[4];
  box[3]=-b[2]; box[4]=-b[3]; box[5]=-b[5];
}

inline void bboxCube(const double *c,int n,double cube[4])
// wolkencanvas.cpp:502-519: BoundRect over the header corners (boundrect.cpp:60-73, orientation 0;
// xy::dirbound point.cpp:85-93 multiplies by the long-double-derived cos/sin of k*90 degrees),
// cube side = largest extent, centre = box middle.
{
  double b[6];
  for (int k=0;k<6;k++)
    b[k]=INFINITY;
  for (int i=0;i<n;i++)
  {
    for (int k=0;k<4;k++)
    {
      int ang=(int)((unsigned)k*0x20000000u);
      double s=(double)sinl(ang*M_PIl/1073741824.),co=(double)cosl(ang*M_PIl/1073741824.);
      double v=c[3*i]*co+c[3*i+1]*s;
      if (v<b[k]) b[k]=v;
    }
    if (c[3*i+2]<b[4]) b[4]=c[3*i+2];
    if (-c[3*i+2]<b[5]) b[5]=-c[3*i+2];
  }
  double left=b[0],bottom=b[1],right=-b[2],top=-b[3],low=b[4],high=-b[5];
  double side=right-left;
  if (top-bottom>side) side=top-bottom;
  if (high-low>side) side=high-low;
  cube[0]=(right+left)/2;
  cube[1]=(top+bottom)/2;
  cube[2]=(high+low)/2;
  cube[3]=side;
}

inline int snakeSetSize(double cubeSide,double tileSize,double *spacing,int *lo,int *hi)
// Flowsnake::setSize, flowsnake.cpp:208-230
{
  int best=0;
  double bestDiff=INFINITY;
  for (int i=0;i<12;i++)
  {
    double diff=fabs(log(cubeSide/kSquareSides[i]/tileSize));
    if (diff<bestDiff)
    {
      bestDiff=diff;
      best=i;
    }
  }
  *spacing=cubeSide/kSquareSides[best];
  *lo=kLoLim[best];
  *hi=kHiLim[best];
  return best;
}

inline std::string ldecimal(double x)
// ldecimal(x,0), ldecimal.cpp:31-127: fewest digits that read back equal, then the layout rules
{
  char buf[64];
  int prec;
  for (prec=0;prec<=DBL_DIG+3;prec++)
  {
    snprintf(buf,sizeof(buf),"%.*e",prec,x);
    if (atof(buf)==x)
      break;
  }
  char *e=strchr(buf,'e');
  int iexp=atoi(e+1);
  *e=0;
  std::string sign,digits;
  for (char *p=buf;*p;p++)
    if (*p=='-')
      sign="-";
    else if (*p!='.')
      digits+=*p;
  while (digits.size()>1 && digits.back()=='0')
    digits.pop_back();
  std::string m=digits.substr(0,1),a=digits.substr(1);
  if (iexp<0 && iexp>-5)
  {
    a=m+a;
    m="";
    iexp++;
  }
  if (iexp>0)
  {
    size_t ch=(size_t)iexp>a.size()?a.size():(size_t)iexp;
    m+=a.substr(0,ch);
    a.erase(0,ch);
    iexp-=(int)ch;
  }
  while (iexp>-5 && iexp<0 && m.empty())
  {
    a="0"+a;
    iexp++;
  }
  while (iexp<3 && iexp>0 && a.empty())
  {
    m+='0';
    iexp--;
  }
  std::string ret=sign+m;
  if (!a.empty())
    ret+="."+a;
  if (iexp)
    ret+="e"+std::to_string(iexp);
  return ret;
}

inline void leafCube(unsigned long long key,int depth,const double rootCenter[3],double rootSide,
                     double center[3],double *half)
// centre of the depth-th cube on the key's path: Octree::cube/split, octree.cpp:335-337, 348-358
{
  double c[3]={rootCenter[0],rootCenter[1],rootCenter[2]},q=rootSide/4;
  for (int l=0;l<depth;l++)
  {
    int d=(int)((key>>(3*(20-l)))&7);
    c[0]+=(d&1)?q:-q;
    c[1]+=(d&2)?q:-q;
    c[2]+=(d&4)?q:-q;
    q/=2;
  }
  center[0]=c[0]; center[1]=c[1]; center[2]=c[2];
  *half=q*2;          // after `depth` steps q = side/2^(depth+2); half side = side/2^(depth+1)
}

} // namespace wbhost
