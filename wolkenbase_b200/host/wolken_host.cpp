// wolken_host.cpp — implementation of the reference-shaped C++ surface over the C ABI.
#include <algorithm>
#include <cassert>
#include <climits>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include "wolken_host.h"

using namespace std;

Octree octRoot;
OctStore octStore;
CloudOutput cloudOutput;
Flowsnake snake;
TileTable tiles;
map<int,size_t> classTotals;
double minHyperboloidSize=0.1,maxSlope=1,thickness=0,tileSize=1;
bool keepRecordsOnDevice=false;
bool hostQueries=false;
double hostTimes[4]={0,0,0,0};        // seconds in wb_create, wb_add_las_file, wb_build, wb_encode+write
ShardReport shardReport;

namespace
{
wb_ctx *g_ctx=nullptr;
int g_nGpus=1;
int g_command=TH_WAIT;
struct InputSeg { LasHeader *h; size_t firstRec,count; };   // a run of records of one open file
vector<InputSeg> g_files;              // in read order: input index = concatenation of these runs
vector<size_t> g_fileFirst;            // input index of each run's first record
// embufferPoint: records waiting to be handed to the device (runs of consecutive readPoint results)
vector<InputSeg> g_pending;
size_t g_pendingPoints=0;
thread_local LasHeader *t_lastHdr=nullptr;      // the most recent readPoint on this thread: header, record, location
thread_local size_t t_lastNum=0;
thread_local double t_lastLoc[3]={0,0,0};
vector<double> g_corners;
Cube g_snakeCube;
double g_snakeTile=1;
bool g_built=false,g_scanned=false,g_postscanned=false,g_classified=false;
// host mirrors, filled on demand
bool g_haveStore=false,g_haveLeaves=false;
vector<wb_leaf> g_leaves;
vector<uint64_t> g_leafLo;             // smallest key of each leaf's cube
vector<double> g_x,g_y,g_z;
vector<uint32_t> g_perm;
vector<uint64_t> g_keys;
vector<uint8_t> g_labels;
vector<uint32_t> g_attrSrc;           // canonical position -> input record whose attributes the stored point has
deque<ThreadAction> g_results;
std::map<uint32_t,uint32_t> g_lastAtLocation;   // duplicates: surviving record -> last record at the same XYZ (valid per build)
bool g_haveLastAtLocation=false;

void flushPointBuffer();

struct Stopwatch
{
  double &acc;
  std::chrono::steady_clock::time_point t0;
  explicit Stopwatch(double &a):acc(a),t0(std::chrono::steady_clock::now()) {}
  ~Stopwatch() { acc+=std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count(); }
};

void die(const char *what)
{
  cerr<<what<<": "<<(g_ctx?wb_last_error(g_ctx):"no context")<<endl;
}

void ensureContext()
{
  if (!g_ctx)
  {
    const char *d=getenv("WOLKEN_DEVICE");
    Stopwatch sw(hostTimes[0]);
    if (wb_create(d?atoi(d):0,&g_ctx)!=WB_OK)
    {
      cerr<<"wolkenbase_b200: no usable CUDA device (there is no CPU path)\n";
      exit(3);
    }
    wb_keep_records(g_ctx,keepRecordsOnDevice?1:0);
  }
}

uint64_t keyOf(xyz p)
// 21 steps of Octree::findBlock's descent, octree.cpp:199-216
{
  xyz c=octRoot.getCenter();
  double cx=c.getx(),cy=c.gety(),cz=c.getz(),q=octRoot.getSide()/4;
  uint64_t key=0;
  for (int l=0;l<WB_LEVELS;l++)
  {
    int xb=p.getx()>=cx,yb=p.gety()>=cy,zb=p.getz()>=cz;
    key=(key<<3)|(uint64_t)(zb*4+yb*2+xb);
    cx+=xb?q:-q; cy+=yb?q:-q; cz+=zb?q:-q;
    q/=2;
  }
  return key;
}

void ensureBuilt()
{
  ensureContext();
  flushPointBuffer();
  if (g_built)
    return;
  double rc[3]={octRoot.getCenter().getx(),octRoot.getCenter().gety(),octRoot.getCenter().getz()};
  double cube[4]={g_snakeCube.getCenter().getx(),g_snakeCube.getCenter().gety(),g_snakeCube.getCenter().getz(),g_snakeCube.getSide()};
  wb_set_params(g_ctx,g_snakeTile,maxSlope,thickness,minHyperboloidSize);
  if (octRoot.getSide()>0 && g_snakeCube.getSide()>0)
    wb_set_geometry(g_ctx,rc,octRoot.getSide(),cube);
  Stopwatch sw(hostTimes[2]);
  if (wb_build(g_ctx)!=WB_OK)
  {
    die("build");
    exit(4);
  }
  g_built=true;
  g_haveStore=g_haveLeaves=g_haveLastAtLocation=false;
}

void ensureLeaves()
// the bucket list alone (enough for getNumBlocks, dump and the device writer)
{
  ensureBuilt();
  if (g_haveLeaves)
    return;
  uint64_t nl=0;
  wb_num_leaves(g_ctx,&nl);
  g_leaves.resize(nl);
  if (nl)
    wb_get_leaves(g_ctx,g_leaves.data(),nl);
  g_haveLeaves=true;
}

void ensureStore()
{
  ensureLeaves();
  if (g_haveStore)
    return;
  uint64_t nl=g_leaves.size();
  wb_stats st;
  wb_get_stats(g_ctx,&st);
  size_t n=st.n_points;
  g_x.resize(n); g_y.resize(n); g_z.resize(n); g_perm.resize(n); g_keys.resize(n);
  wb_get_points_sorted(g_ctx,g_x.data(),g_y.data(),g_z.data());
  wb_get_order(g_ctx,g_perm.data(),g_keys.data());
  g_leafLo.resize(nl);
  for (size_t i=0;i<nl;i++)
  {
    int s=3*(WB_LEVELS-g_leaves[i].depth);
    g_leafLo[i]=s>=63?0:(g_keys[g_leaves[i].first]>>s)<<s;
  }
  // OctBuffer::put overwrites the stored point with a newcomer at the same XYZ (octree.cpp:626-644):
  // the place is the first record's, the attributes the last one's
  g_attrSrc=g_perm;
  if (st.n_duplicates)
  {
    vector<uint32_t> dup(st.n_duplicates),rep(st.n_duplicates);
    wb_get_duplicates(g_ctx,dup.data(),rep.data(),st.n_duplicates);
    std::map<uint32_t,uint32_t> last;
    for (size_t u=0;u<dup.size();u++)
    {
      uint32_t &l=last[rep[u]];
      if (dup[u]>l)
        l=dup[u];
    }
    for (size_t k=0;k<n;k++)
    {
      auto it=last.find(g_perm[k]);
      if (it!=last.end())
        g_attrSrc[k]=it->second;
    }
  }
  g_haveStore=true;
}

void ensureLabels()
{
  if (!g_classified)
    return;
  size_t n=0;
  for (auto &f:g_files)
    n+=f.count;
  if (g_labels.size()!=n)
  {
    g_labels.resize(n);
    wb_get_labels(g_ctx,g_labels.data());
  }
}

LasPoint pointAt(size_t k)
// k-th point of the canonical order, all attributes decoded from its original record
{
  uint32_t i=g_attrSrc[k];
  size_t f=upper_bound(g_fileFirst.begin(),g_fileFirst.end(),(size_t)i)-g_fileFirst.begin()-1;
  LasPoint p=g_files[f].h->readPoint(g_files[f].firstRec+(i-g_fileFirst[f]));
  if (p.returnNum==0)
    p.returnNum=1;                     // a stored point with return number 0: keep-zeros file, threads.cpp:527-528
  p.location=xyz(g_x[k],g_y[k],g_z[k]);
  if (g_classified)
  {
    ensureLabels();
    p.classification=g_labels[i];
  }
  return p;
}

Cube leafCube(size_t i)
{
  return Cube(xyz(g_leaves[i].cx,g_leaves[i].cy,g_leaves[i].cz),2*g_leaves[i].half);
}

void findBlocksRec(const Shape &sh,size_t lo,size_t hi,int depth,xyz c,double side,vector<int64_t> &out)
// Octree::findBlocks (octree.cpp:234-251) over the leaf list: [lo,hi) are the leaves below the
// node of centre c and side `side` at `depth`
{
  if (!sh.intersect(Cube(c,side)))
    return;
  for (int ch=0;ch<8;ch++)
  {
    int s=3*(WB_LEVELS-1-depth);
    size_t a=lo;
    while (a<hi && (int)((g_leafLo[a]>>s)&7)<ch) a++;
    size_t b=a;
    while (b<hi && (int)((g_leafLo[b]>>s)&7)==ch) b++;
    if (a==b)
      continue;
    xyz cc(c.getx()+((ch&1)?side/4:-side/4),c.gety()+((ch&2)?side/4:-side/4),c.getz()+((ch&4)?side/4:-side/4));
    if (b-a==1 && g_leaves[a].depth==depth+1)
    {
      if (sh.intersect(Cube(cc,side/2)))
        out.push_back((int64_t)a);
    }
    else
      findBlocksRec(sh,a,b,depth+1,cc,side/2,out);
    lo=b;
  }
}
} // namespace

// ---------------------------------------------------------------- shapes (shape.cpp)
bool Cube::in(xyz p) const
{
  return fabs(p.getx()-center.getx())<=side/2 && fabs(p.gety()-center.gety())<=side/2 && fabs(p.getz()-center.getz())<=side/2;
}

xyz Cube::corner(int n) const
{
  return xyz(center.getx()+((n&1)?side/2:-side/2),center.gety()+((n&2)?side/2:-side/2),center.getz()+((n&4)?side/2:-side/2));
}

bool Shape::in(Cube &cube) const
{
  bool ret=true;
  for (int i=0;i<8;i++)
    ret=ret && in(cube.corner(i));
  return ret;
}

static double clampToward(double v,double c,double half)
// the coordinate of the cube (centre c) nearest to v
{
  if (fabs(v-c)<half)
    return v;
  return v>c?c+half:c-half;
}

bool Paraboloid::in(xyz p) const
{
  double d=dist(xy(vertex),xy(p)),zd=vertex.getz()-p.getz();
  if (radiusCurvature==0)
    return d==0;
  return 2*zd/radiusCurvature>=sqr(d/radiusCurvature);
}

xyz Paraboloid::closestPoint(Cube cube) const
{
  xyz c=cube.getCenter();
  double h=cube.getSide()/2,z=c.getz();
  if (radiusCurvature>0) z-=h;
  if (radiusCurvature<0) z+=h;
  return xyz(clampToward(vertex.getx(),c.getx(),h),clampToward(vertex.gety(),c.gety(),h),z);
}

Hyperboloid::Hyperboloid(xyz v,double r,double s)
{
  double por=r*sqr(s);
  slope=s;
  por2=sqr(por);
  center=v+xyz(0,0,por);
  vertex0=v;
  r0=r;
}

static bool fillShape(wb_shape &s,int type,double a,double b,double c,double d,double e)
{
  memset(&s,0,sizeof(s));
  s.type=type;
  s.p[0]=a; s.p[1]=b; s.p[2]=c; s.p[3]=d; s.p[4]=e;
  return true;
}
bool Paraboloid::abi(wb_shape &s) const { return fillShape(s,WB_PARABOLOID,vertex.getx(),vertex.gety(),vertex.getz(),radiusCurvature,0); }
bool Hyperboloid::abi(wb_shape &s) const { return fillShape(s,WB_HYPERBOLOID,vertex0.getx(),vertex0.gety(),vertex0.getz(),r0,slope); }
bool Sphere::abi(wb_shape &s) const { return fillShape(s,WB_SPHERE,center.getx(),center.gety(),center.getz(),radius,0); }
bool Cylinder::abi(wb_shape &s) const { return fillShape(s,WB_CYLINDER,center.getx(),center.gety(),radius,0,0); }
bool Column::abi(wb_shape &s) const { return fillShape(s,WB_COLUMN,center.getx(),center.gety(),side,0,0); }

bool Column::in(xyz p) const
{
  return fabs(center.getx()-p.getx())<=side/2 && fabs(center.gety()-p.gety())<=side/2;
}

xyz Column::closestPoint(Cube cube) const
{
  xyz c=cube.getCenter();
  double h=cube.getSide()/2;
  return xyz(clampToward(center.getx(),c.getx(),h),clampToward(center.gety(),c.gety(),h),c.getz());
}

bool Hyperboloid::in(xyz p) const
{
  double d=dist(xy(center),xy(p)),zd=center.getz()-p.getz();
  if (slope>0)
    return zd>0 && sqr(zd)-sqr(d*slope)>=por2;
  return zd<0 && sqr(zd)-sqr(d*slope)>=por2;
}

xyz Hyperboloid::closestPoint(Cube cube) const
{
  xyz c=cube.getCenter();
  double h=cube.getSide()/2,z=c.getz();
  if (slope>0) z-=h;
  if (slope<0) z+=h;
  return xyz(clampToward(center.getx(),c.getx(),h),clampToward(center.gety(),c.gety(),h),z);
}

bool Sphere::in(xyz p) const
{
  return hypot(hypot(p.getx()-center.getx(),p.gety()-center.gety()),p.getz()-center.getz())<=radius;
}

xyz Sphere::closestPoint(Cube cube) const
{
  xyz c=cube.getCenter();
  double h=cube.getSide()/2;
  return xyz(clampToward(center.getx(),c.getx(),h),clampToward(center.gety(),c.gety(),h),clampToward(center.getz(),c.getz(),h));
}

bool Cylinder::in(xyz p) const
{
  return dist(center,xy(p))<=radius;
}

xyz Cylinder::closestPoint(Cube cube) const
{
  xyz c=cube.getCenter();
  double h=cube.getSide()/2;
  return xyz(clampToward(center.getx(),c.getx(),h),clampToward(center.gety(),c.gety(),h),c.getz());
}

// ---------------------------------------------------------------- LAS
LasPoint::LasPoint()
{
  location=xyz(NAN,NAN,NAN);
  intensity=returnNum=nReturns=classification=classificationFlags=0;
  scannerChannel=userData=pointSource=nir=red=green=blue=0;
  scanDirection=edgeLine=false;
  scanAngle=0;
  gpsTime=0;
}

LasHeader::LasHeader()
{
  map=nullptr;
  mapLen=0;
  versionMajor=versionMinor=0;
  headerSize=pointOffset=0;
  pointFormat=pointLength=0;
  unit=1;
  zipFlag=false;
  out=nullptr;
  writePos=0;
  for (auto &v:nPoints)
    v=0;
  xScale=yScale=zScale=xOffset=yOffset=zOffset=maxX=minX=maxY=minY=maxZ=minZ=0;
}

LasHeader::LasHeader(const LasHeader &o)
// member-wise, except that a copy never owns the mapping or the output stream
{
  filename=o.filename;
  systemId=o.systemId;
  map=nullptr;
  mapLen=0;
  out=nullptr;
  versionMajor=o.versionMajor; versionMinor=o.versionMinor;
  headerSize=o.headerSize; pointOffset=o.pointOffset;
  pointFormat=o.pointFormat; pointLength=o.pointLength;
  xScale=o.xScale; yScale=o.yScale; zScale=o.zScale;
  xOffset=o.xOffset; yOffset=o.yOffset; zOffset=o.zOffset;
  maxX=o.maxX; minX=o.minX; maxY=o.maxY; minY=o.minY; maxZ=o.maxZ; minZ=o.minZ;
  unit=o.unit;
  memcpy(nPoints,o.nPoints,sizeof(nPoints));
  zipFlag=o.zipFlag;
  writePos=o.writePos;
}

LasHeader::~LasHeader()
{
  close();
}

void LasHeader::close()
{
  if (map)
    munmap(map,mapLen);
  map=nullptr;
  mapLen=0;
  if (out)
    fclose(out);
  out=nullptr;
}

template <typename T> static T rd(const uint8_t *p)
{
  T v;
  memcpy(&v,p,sizeof(T));
  return v;
}

void LasHeader::openRead(string fileName)
// Field layout and the "which point count is right" vote of las.cpp:299-428
{
  close();
  filename=fileName;
  versionMajor=versionMinor=0;
  nPoints[0]=0;
  int fd=open(fileName.c_str(),O_RDONLY);
  if (fd<0)
    return;
  struct stat st;
  if (fstat(fd,&st) || st.st_size<227)
  {
    ::close(fd);
    return;
  }
  mapLen=st.st_size;
  map=(uint8_t *)mmap(nullptr,mapLen,PROT_READ,MAP_PRIVATE,fd,0);
  ::close(fd);
  if (map==MAP_FAILED)
  {
    map=nullptr;
    return;
  }
  if (memcmp(map,"LASF",4))
    return;
  const uint8_t *h=map;
  versionMajor=h[24];
  versionMinor=h[25];
  headerSize=rd<uint16_t>(h+94);
  pointOffset=rd<uint32_t>(h+96);
  pointFormat=h[104];
  pointLength=rd<uint16_t>(h+105);
  unsigned legacy[6];
  for (int i=0;i<6;i++)
    legacy[i]=rd<uint32_t>(h+107+4*i);
  xScale=rd<double>(h+131); yScale=rd<double>(h+139); zScale=rd<double>(h+147);
  xOffset=rd<double>(h+155); yOffset=rd<double>(h+163); zOffset=rd<double>(h+171);
  maxX=rd<double>(h+179); minX=rd<double>(h+187); maxY=rd<double>(h+195);
  minY=rd<double>(h+203); maxZ=rd<double>(h+211); minZ=rd<double>(h+219);
  for (auto &v:nPoints)
    v=0;
  if (headerSize>0xe3 && mapLen>=375)
    for (int i=0;i<16;i++)
      nPoints[i]=rd<uint64_t>(h+247+8*i);
  int which=15;
  size_t total=0;
  for (int i=1;i<6;i++)
  {
    total+=legacy[i];
    if (legacy[i]>legacy[0])
      which&=~5;
  }
  if (total!=legacy[0]) which&=~1;
  if (total!=0 || legacy[0]==0) which&=~4;
  total=0;
  for (int i=1;i<16;i++)
  {
    total+=nPoints[i];
    if (nPoints[i]>nPoints[0])
      which&=~10;
  }
  if (total!=nPoints[0]) which&=~2;
  if (total!=0 || nPoints[0]==0) which&=~8;
  for (int i=0;i<6;i++)
  {
    if (legacy[i]!=nPoints[i] && legacy[i]!=0) which&=~10;
    if (nPoints[i]!=legacy[i] && nPoints[i]!=0) which&=~5;
  }
  if (which>0 && (which&3)==0)
  {
    cerr<<"Number of points by return are all 0. Setting first return to all points.\n";
    nPoints[1]=nPoints[0];
    legacy[1]=legacy[0];
  }
  if (which==1 || which==4)
    for (int i=0;i<6;i++)
      nPoints[i]=legacy[i];
  if (which==0)
    for (int i=0;i<6;i++)
      nPoints[i]=0;
  if (pointLength==0)
    versionMajor=versionMinor=nPoints[0]=0;
  {
    // truncated file: read what is there.  The limit is a quotient (a product of header fields could wrap), and a
    // point offset beyond the file leaves no points at all (advice, round 1: the unsigned difference was huge)
    const uint64_t room=(uint64_t)pointOffset<mapLen?mapLen-(uint64_t)pointOffset:0;
    const uint64_t fit=pointLength?room/pointLength:0;
    if ((uint64_t)nPoints[0]>fit)
      nPoints[0]=fit;
  }
  zipFlag=pointOffset>=headerSize+8 && mapLen>=headerSize+8 && !memcmp(map+headerSize+2,"laszip",6);
  if (zipFlag)
    cout<<filename<<" is laszipped\n";
}

bool LasHeader::isValid() const
{
  static const int minLen[11]={20,28,26,34,57,63,30,36,38,59,67};       // las.cpp:38
  if (pointFormat>=0 && pointFormat<=10 && pointLength<minLen[pointFormat])
    return false;                                    // readPoint would walk past the record
  return versionMajor>0 && versionMinor>0 && headerSize>0 && pointLength>0 && (headerSize>0xe3 || pointFormat<6);
}

static int degToBin(double deg)
// degtobin -> rottobin, angle.cpp:233-251
{
  double ip=0,fp=2*modf(deg/360/2,&ip);
  if (fp>=1) fp-=2;
  if (fp<-1) fp+=2;
  return (int)lrint(2147483648.*fp);
}

LasPoint LasHeader::readPoint(size_t num)
// las.cpp:735-820
{
  LasPoint ret;
  if (!map || num>=nPoints[0])
    throw -1;
  const uint8_t *r=map+pointOffset+num*pointLength;
  int xi=rd<int32_t>(r),yi=rd<int32_t>(r+4),zi=rd<int32_t>(r+8),o;
  ret.intensity=rd<uint16_t>(r+12);
  if (pointFormat<6)
  {
    ret.returnNum=r[14]&7;
    ret.nReturns=(r[14]>>3)&7;
    ret.scanDirection=(r[14]>>6)&1;
    ret.edgeLine=(r[14]>>7)&1;
    ret.classification=r[15]&31;
    ret.classificationFlags=(r[15]>>5)&7;
    ret.scanAngle=degToBin((signed char)r[16]);
    ret.userData=r[17];
    ret.pointSource=rd<uint16_t>(r+18);
    o=20;
  }
  else
  {
    ret.returnNum=r[14]&15;
    ret.nReturns=(r[14]>>4)&15;
    ret.classificationFlags=r[15]&15;
    ret.scannerChannel=(r[15]>>4)&3;
    ret.scanDirection=(r[15]>>6)&1;
    ret.edgeLine=(r[15]>>7)&1;
    ret.classification=r[16];
    ret.userData=r[17];
    ret.scanAngle=degToBin(rd<int16_t>(r+18)*0.006);
    ret.pointSource=rd<uint16_t>(r+20);
    o=22;
  }
  if ((1<<pointFormat)&0x7fa)
  {
    ret.gpsTime=rd<double>(r+o);
    o+=8;
  }
  if ((1<<pointFormat)&0x5ac)
  {
    ret.red=rd<uint16_t>(r+o); ret.green=rd<uint16_t>(r+o+2); ret.blue=rd<uint16_t>(r+o+4);
    o+=6;
  }
  if ((1<<pointFormat)&0x500)
    ret.nir=rd<uint16_t>(r+o);
  ret.location=xyz(xOffset+xScale*xi,yOffset+yScale*yi,zOffset+zScale*zi)*unit;
  t_lastHdr=this;                       // embufferPoint hands this very record to the device
  t_lastNum=num;
  t_lastLoc[0]=ret.location.getx(); t_lastLoc[1]=ret.location.gety(); t_lastLoc[2]=ret.location.getz();
  if (ret.location.getx()>maxX*unit || ret.location.getx()<minX*unit || ret.location.gety()>maxY*unit ||
      ret.location.gety()<minY*unit || ret.location.getz()>maxZ*unit || ret.location.getz()<minZ*unit)
    cerr<<"Point out of range\n";
  return ret;
}

// ---------------------------------------------------------------- flowsnake / tiles
Eisenstein toFlowsnake(int n)
{
  static const unsigned char tab[6][7]=
  {
    {0x52,0x05,0x06,0x24,0x33,0x40,0x01},{0x31,0x10,0x12,0x05,0x43,0x54,0x16},{0x46,0x24,0x21,0x10,0x53,0x35,0x22},
    {0x31,0x10,0x03,0x54,0x36,0x35,0x22},{0x46,0x24,0x13,0x35,0x42,0x40,0x01},{0x52,0x05,0x23,0x40,0x51,0x54,0x16}
  };
  int dig[11],ori=0;
  long long v=(long long)n+1235829214LL;
  for (int i=0;i<11;i++) { dig[i]=(int)(v%7); v/=7; }
  for (int i=10;i>=0;i--) { int t=tab[ori][dig[i]]; ori=t>>4; dig[i]=t&7; }
  int x=0,y=0,px=1,py=0;
  for (int i=0;i<11;i++)
  {
    int d=dig[i]-3,dy=(d+4)/3-1,dx=d-2*dy;
    x+=dx*px-dy*py;
    y+=dx*py+dy*px-dy*py;
    int nx=2*px+py,ny=3*py-px;
    px=nx; py=ny;
  }
  return Eisenstein(x,y);
}

void Flowsnake::setSize(Cube cube,double desiredSpacing)
{
  int lo,hi;
  assert(desiredSpacing>0 && cube.getSide()>0);
  wb_snake_set_size(cube.getSide(),desiredSpacing,&spacing,&lo,&hi);
  center=xy(cube.getCenter());
  startnum=counter=lo;
  stopnum=hi;
  g_snakeCube=cube;
  g_snakeTile=desiredSpacing;
  g_built=g_scanned=g_postscanned=g_classified=false;
}

Eisenstein Flowsnake::next()
{
  if (counter<=stopnum)
    return toFlowsnake(counter++);
  return Eisenstein(INT_MIN,INT_MIN);
}

Cylinder Flowsnake::cyl(Eisenstein e)
{
  double rad=spacing*41/71;
  double re=(e.getx()-e.gety()/2.)*spacing,im=e.gety()*0.86602540378443864676372317*spacing;
  if (e.getx()==INT_MIN)
    rad=0;
  return Cylinder(xy(re,im)+center,rad);
}

void TileTable::sync()
{
  if (stale)
  {
    stale=false;
    refreshTiles();
  }
}

Tile &TileTable::operator[](Eisenstein e)
{
  sync();
  auto it=byAddr.find(e);
  if (it==byAddr.end())
  {
    Tile z;
    memset(&z,0,sizeof(z));
    it=byAddr.insert(make_pair(e,z)).first;
  }
  return it->second;
}

void refreshTiles()
{
  uint64_t n=0;
  tiles.byAddr.clear();
  if (wb_num_tiles(g_ctx,&n)!=WB_OK || !n)
    return;
  vector<wb_tile> t(n);
  wb_get_tiles(g_ctx,t.data(),n);
  for (auto &w:t)
  {
    Tile z;
    memset(&z,0,sizeof(z));
    z.nPoints=w.nPoints;
    z.treeFlags=(short)w.treeFlags;
    z.density=w.density;
    z.hyperboloidSize=w.hyperboloidSize;
    z.height=w.height;
    tiles.byAddr[Eisenstein(w.ex,w.ey)]=z;
  }
}

void initTiles()
{
  tiles.clear();
}

// ---------------------------------------------------------------- octree / store
void Octree::sizeFit(vector<xyz> pnts)
{
  vector<double> c;
  for (auto &p:pnts)
  {
    c.push_back(p.getx()); c.push_back(p.gety()); c.push_back(p.getz());
  }
  double ctr[3];
  wb_size_fit(c.data(),(int)pnts.size(),ctr,&side);
  center=xyz(ctr[0],ctr[1],ctr[2]);
  g_corners=c;
  g_built=false;
}

void Octree::clear()
{
  side=0;
}

int64_t Octree::findBlock(xyz pnt)
{
  ensureStore();
  if (g_leaves.empty())
    return -1;
  uint64_t k=keyOf(pnt);
  size_t i=upper_bound(g_leafLo.begin(),g_leafLo.end(),k)-g_leafLo.begin();
  if (!i)
    return -1;
  i--;
  int s=3*(WB_LEVELS-g_leaves[i].depth);
  return (k>>s)==(g_leafLo[i]>>s)?(int64_t)i:-1;
}

Cube Octree::findCube(xyz pnt)
{
  int64_t b=findBlock(pnt);
  if (b>=0)
    return leafCube(b);
  // an empty octant: descend as far as the populated tree goes (octree.cpp:218-232)
  xyz c=center;
  double s=side;
  uint64_t k=keyOf(pnt);
  size_t lo=0,hi=g_leaves.size();
  for (int depth=0;depth<WB_LEVELS;depth++)
  {
    int ch=(int)((k>>(3*(WB_LEVELS-1-depth)))&7),sh=3*(WB_LEVELS-1-depth);
    c=xyz(c.getx()+((ch&1)?s/4:-s/4),c.gety()+((ch&2)?s/4:-s/4),c.getz()+((ch&4)?s/4:-s/4));
    s/=2;
    size_t a=lo;
    while (a<hi && (int)((g_leafLo[a]>>sh)&7)<ch) a++;
    size_t b2=a;
    while (b2<hi && (int)((g_leafLo[b2]>>sh)&7)==ch) b2++;
    if (a==b2)
      break;
    lo=a; hi=b2;
  }
  return Cube(c,s);
}

vector<int64_t> Octree::findBlocks(const Shape &sh)
{
  vector<int64_t> out;
  ensureStore();
  if (!g_leaves.empty())
    findBlocksRec(sh,0,g_leaves.size(),0,center,side,out);
  return out;
}

size_t OctStore::getNumBlocks()
{
  ensureLeaves();
  return g_leaves.size();
}

vector<LasPoint> OctStore::getAll(int64_t block)
{
  vector<LasPoint> ret;
  ensureStore();
  if (block<0 || (size_t)block>=g_leaves.size())
    return ret;
  for (size_t k=g_leaves[block].first;k<g_leaves[block].first+g_leaves[block].count;k++)
    ret.push_back(pointAt(k));
  return ret;
}

static bool lowerThan(const LasPoint &a,const LasPoint &b)
{
  return a.location.getz()<b.location.getz();
}

static LasPoint pointFromInput(uint32_t inputIdx,double x,double y,double z)
// the stored point at (x,y,z) whose first record is inputIdx, without the host copy of the store
{
  std::map<uint32_t,uint32_t> &last=g_lastAtLocation;   // survivor -> last record at its location
  if (!g_haveLastAtLocation)
  {
    last.clear();
    wb_stats st;
    wb_get_stats(g_ctx,&st);
    if (st.n_duplicates)
    {
      vector<uint32_t> dup(st.n_duplicates),rep(st.n_duplicates);
      wb_get_duplicates(g_ctx,dup.data(),rep.data(),st.n_duplicates);
      for (size_t u=0;u<dup.size();u++)
      {
        uint32_t &l=last[rep[u]];
        if (dup[u]>l)
          l=dup[u];
      }
    }
    g_haveLastAtLocation=true;
  }
  uint32_t i=inputIdx;
  auto it=last.find(i);
  if (it!=last.end())
    i=it->second;
  size_t f=upper_bound(g_fileFirst.begin(),g_fileFirst.end(),(size_t)i)-g_fileFirst.begin()-1;
  LasPoint p=g_files[f].h->readPoint(g_files[f].firstRec+(i-g_fileFirst[f]));
  if (p.returnNum==0)
    p.returnNum=1;
  p.location=xyz(x,y,z);
  if (g_classified)
  {
    ensureLabels();
    p.classification=g_labels[inputIdx];
  }
  return p;
}

vector<LasPoint> OctStore::pointsIn(const Shape &sh,bool sorted)
{
  vector<LasPoint> ret;
  wb_shape ws;
  if (!hostQueries && sh.abi(ws))
  { // on the device: exact predicate over the whole store, results in bucket order
    ensureBuilt();
    uint64_t n=0;
    if (wb_query_points(g_ctx,&ws,0,&n,nullptr,nullptr,nullptr,nullptr,nullptr)!=WB_OK)
    {
      die("query");
      return ret;
    }
    vector<uint32_t> idx(n);
    vector<double> x(n),y(n),z(n);
    if (n && wb_query_points(g_ctx,&ws,n,&n,nullptr,idx.data(),x.data(),y.data(),z.data())!=WB_OK)
    {
      die("query");
      return ret;
    }
    ret.reserve(n);
    for (size_t k=0;k<n;k++)
      ret.push_back(pointFromInput(idx[k],x[k],y[k],z[k]));
    if (sorted)
      sort(ret.begin(),ret.end(),lowerThan);
    return ret;
  }
  ensureStore();
  for (int64_t b:octRoot.findBlocks(sh))
  {
    Cube cube=leafCube(b);
    bool all=sh.in(cube);
    for (size_t k=g_leaves[b].first;k<g_leaves[b].first+g_leaves[b].count;k++)
      if (all || sh.in(xyz(g_x[k],g_y[k],g_z[k])))
        ret.push_back(pointAt(k));
  }
  if (sorted)
    sort(ret.begin(),ret.end(),lowerThan);
  return ret;
}

uint64_t OctStore::countPointsIn(const Shape &sh)
{
  uint64_t ret=0;
  wb_shape ws;
  if (!hostQueries && sh.abi(ws))
  {
    ensureBuilt();
    if (wb_query_batch(g_ctx,&ws,1,&ret,nullptr,nullptr)!=WB_OK)
      die("query");
    return ret;
  }
  ensureStore();
  for (int64_t b:octRoot.findBlocks(sh))
  {
    Cube cube=leafCube(b);
    if (sh.in(cube))
      ret+=g_leaves[b].count;
    else
      for (size_t k=g_leaves[b].first;k<g_leaves[b].first+g_leaves[b].count;k++)
        ret+=sh.in(xyz(g_x[k],g_y[k],g_z[k]));
  }
  return ret;
}

array<double,2> OctStore::hiLoPointsIn(const Shape &sh)
{
  double hi=-INFINITY,lo=INFINITY;
  wb_shape ws;
  if (!hostQueries && sh.abi(ws))
  {
    ensureBuilt();
    if (wb_query_batch(g_ctx,&ws,1,nullptr,&lo,&hi)!=WB_OK)
      die("query");
    return array<double,2>{lo,hi};
  }
  ensureStore();
  for (int64_t b:octRoot.findBlocks(sh))
    if (g_leaves[b].low<lo || g_leaves[b].high>hi)
      for (size_t k=g_leaves[b].first;k<g_leaves[b].first+g_leaves[b].count;k++)
        if (sh.in(xyz(g_x[k],g_y[k],g_z[k])))
        {
          hi=max(hi,g_z[k]);
          lo=min(lo,g_z[k]);
        }
  return array<double,2>{lo,hi};
}

map<int,size_t> OctStore::countClasses(int64_t block)
{
  map<int,size_t> ret;
  for (auto &p:getAll(block))
    ret[p.classification]++;
  return ret;
}

void OctStore::dump(ofstream &file)
{
  ensureLeaves();
  vector<char> buf(g_leaves.size()*128+64);
  int n=wb_format_dump(g_leaves.data(),g_leaves.size(),buf.data(),buf.size());
  if (n>0)
    file.write(buf.data(),n);
}

uint64_t OctStore::countPoints()
{
  ensureStore();
  return g_x.size();
}

void OctStore::clear()
{
  if (g_ctx)
    wb_clear(g_ctx);
  g_files.clear();
  g_fileFirst.clear();
  g_pending.clear();
  g_pendingPoints=0;
  g_built=g_scanned=g_postscanned=g_classified=g_haveStore=g_haveLeaves=false;
  g_labels.clear();
}

// ---------------------------------------------------------------- cloud.cpp
vector<xyz> cloud;

int64_t getNumCloudBlocks()
{
  return (int64_t)((cloud.size()+RECORDS-1)/RECORDS);
}

vector<LasPoint> getCloudBlock(int64_t n)
{
  vector<LasPoint> ret;
  if (n<0)
    return ret;
  LasPoint pnt;
  for (size_t i=(size_t)n*RECORDS;i<(size_t)(n+1)*RECORDS && i<cloud.size();i++)
  {
    pnt.location=cloud[i];
    ret.push_back(pnt);
  }
  return ret;
}

// ---------------------------------------------------------------- testpattern.cpp: the census
namespace
{
vector<uint64_t> g_pointCensus;
}

int censusPoints(vector<LasPoint> points)
// 1: a number was already counted; -1: a GPS time is no point number; 0: all new
{
  int ret=0;
  for (size_t i=0;i<points.size() && ret>=0;i++)
  {
    const double t=points[i].gpsTime;
    if (!(t>=0) || t>=2147483648.0 || t!=rint(t))
      ret=-1;
    else
    {
      const uint64_t n=(uint64_t)t;
      if (g_pointCensus.size()<n/64+1)
        g_pointCensus.resize(n/64+1);
      const uint64_t mask=(uint64_t)1<<(n%64);
      if (g_pointCensus[n/64]&mask)
        ret=1;
      g_pointCensus[n/64]|=mask;
    }
  }
  return ret;
}

long long censusPoints(ostream *report)
{
  ostream &os=report?*report:cout;
  ensureBuilt();
  int err=0;
  uint64_t maxPoint=0,nMissing=0;
  vector<uint64_t> missing;
  if (keepRecordsOnDevice)
  {
    wb_census_result r;
    missing.resize(4096);
    if (wb_census(g_ctx,&r,missing.data(),missing.size())!=WB_OK)
    {
      die("census");
      return -1;
    }
    err=r.status;
    maxPoint=r.max_point;
    nMissing=r.n_missing;
    missing.resize(min<uint64_t>(missing.size(),nMissing));
  }
  else
  {
    g_pointCensus.clear();
    bool dup=false;
    for (size_t i=0;i<octStore.getNumBlocks() && err>=0;i++)
    {
      err=censusPoints(octStore.getAll((int64_t)i));
      dup=dup || err>0;
    }
    if (err>=0)
    {
      err=dup?1:0;
      int b=0;
      while (b<64 && !g_pointCensus.empty() && (g_pointCensus.back()>>b))
        b++;
      maxPoint=g_pointCensus.empty()?0:(g_pointCensus.size()-1)*64+b;
      for (uint64_t i=0;i<maxPoint;i++)
      {
        if (g_pointCensus[i/64]==~(uint64_t)0)
        {
          i+=63-(i&63);
          continue;
        }
        if (!(g_pointCensus[i/64]&((uint64_t)1<<(i&63))))   // the reference shifts an int here (testpattern.cpp:108)
        {
          nMissing++;
          if (missing.size()<4096)
            missing.push_back(i);
        }
      }
    }
  }
  if (err<0)
    return -1;
  if (err>0 && maxPoint>64)                          // all GPS times 0: the format may simply have none
    os<<"Duplicate point\n";
  os<<"Max point "<<maxPoint<<endl;
  if (nMissing)
  {
    os<<"Missing points: ";
    for (size_t i=0;i<missing.size();i++)
      os<<(i?",":"")<<missing[i];
    if (nMissing>missing.size())
      os<<",... ("<<nMissing<<" in all)";
    os<<endl;
  }
  return (long long)nMissing;
}

// ---------------------------------------------------------------- the phase protocol (threads.h)
void startThreads(int n)
// threads.cpp:91-113 starts n worker threads; here n is the number of GPUs the tile phases are spread over (one
// worker = one GPU with its own context and its x-strip of the cloud, csrc/wb_shard.cuh).  The store that answers
// queries and feeds the writer stays on the first device.
{
  ensureContext();
  int ndev=1;
  wb_device_count(&ndev);
  const char *share=getenv("WOLKEN_TRANSPORT");
  g_nGpus=max(1,n);
  if (!(share && !strcmp(share,"local")))          // NCCL wants one device per rank; the LOCAL transport can share one
    g_nGpus=min(g_nGpus,max(1,ndev));
  g_command=TH_WAIT;
}

void joinThreads() {}
int nThreads() { return g_nGpus; }
double busyFraction() { return 0; }
bool actionQueueEmpty() { return true; }
bool resultQueueEmpty() { return g_results.empty(); }
bool pointBufferEmpty() { return g_pendingPoints==0; }
size_t pointBufferSize() { return g_pendingPoints; }

size_t duplicatePoints()
{
  ensureBuilt();
  wb_stats st;
  wb_get_stats(g_ctx,&st);
  return st.n_duplicates;
}
void setThreadCommand(int s) { waitForThreads(s); }
int getThreadCommand() { return g_command; }
int getThreadStatus() { return (g_command<<20)|g_command; }
void fillTanTables() {}

ThreadAction dequeueResult()
{
  ThreadAction a;
  if (!g_results.empty())
  {
    a=g_results.front();
    g_results.pop_front();
  }
  return a;
}

namespace
{
void sendExtents()
{
  if (g_files.empty())
    for (size_t i=0;i+5<g_corners.size();i+=6)
      wb_add_extent(g_ctx,&g_corners[i],&g_corners[i+3]);
}

void addInputRun(LasHeader *h,size_t firstRec,size_t count)
{
  size_t first=g_fileFirst.empty()?0:g_fileFirst.back()+g_files.back().count;
  g_files.push_back(InputSeg{h,firstRec,count});
  g_fileFirst.push_back(first);
  g_built=g_scanned=g_postscanned=g_classified=false;
}

void flushPointBuffer()
// The embuffered records go to the device as they lie in their files (runs of consecutive records).
{
  if (g_pending.empty())
    return;
  ensureContext();
  wb_set_return_zero_rule(g_ctx,1);     // the caller chose the points: none is dropped (wolkencli.cpp:104-108)
  for (auto &r:g_pending)
  {
    LasHeader *h=r.h;
    sendExtents();
    double sc[3]={h->rawScale(0),h->rawScale(1),h->rawScale(2)},of[3]={h->rawOffset(0),h->rawOffset(1),h->rawOffset(2)};
    Stopwatch sw(hostTimes[1]);
    if (wb_add_las(g_ctx,h->records()+r.firstRec*h->getPointLength(),r.count,h->getPointFormat(),h->getPointLength(),
                   sc,of,h->getUnit())!=WB_OK)
    {
      die("Error storing points");
      break;
    }
    addInputRun(h,r.firstRec,r.count);
  }
  wb_set_return_zero_rule(g_ctx,0);
  g_pending.clear();
  g_pendingPoints=0;
}
} // namespace

void embufferPoint(LasPoint point,bool fromFile)
// threads.cpp:201-229.  The device store keeps the file's own integers, so the point must be the one the last
// readPoint on this thread returned (the reference's callers, wolkencli.cpp:104-108 and threads.cpp:525-531,
// do exactly that); anything else is refused aloud.
{
  (void)fromFile;
  if (point.isEmpty())
    return;
  if (!t_lastHdr || point.location.getx()!=t_lastLoc[0] || point.location.gety()!=t_lastLoc[1] ||
      point.location.getz()!=t_lastLoc[2])
  {
    cerr<<"embufferPoint: not the point readPoint just returned; ignored (the GPU store holds LAS records)\n";
    return;
  }
  if (!g_pending.empty() && g_pending.back().h==t_lastHdr && g_pending.back().firstRec+g_pending.back().count==t_lastNum)
    g_pending.back().count++;
  else
    g_pending.push_back(InputSeg{t_lastHdr,t_lastNum,1});
  g_pendingPoints++;
}

void embufferPoints(vector<LasPoint> points,int)
// threads.cpp:231-246 re-embuffers the 537 points of a split block; there are no such splits here.
{
  if (!points.empty())
    cerr<<"embufferPoints: nothing to re-insert, buckets are split on the device\n";
}

LasPoint debufferPoint(int) { return LasPoint(); }   // empty point = "buffer is empty", threads.cpp:248-264
void sleepDead(int) {}
int thisThread() { return -1; }                      // the caller's thread, threads.cpp:414-417
bool tileDoneQueueEmpty() { return true; }
Eisenstein dequeueTileDone() { return Eisenstein(INT_MIN,INT_MIN); }

void enqueueAction(ThreadAction a)
{
  ensureContext();
  switch (a.opcode)
  {
    case ACT_READ:
    {
      LasHeader *h=a.hdr;
      if (!h || !h->isValid() || h->isZipped())
      {
        cerr<<"Error reading file\n";
        break;
      }
      cout<<"Thread 0 reading "<<h->getFileName()<<endl;
      flushPointBuffer();
      sendExtents();
      double sc[3]={h->rawScale(0),h->rawScale(1),h->rawScale(2)},of[3]={h->rawOffset(0),h->rawOffset(1),h->rawOffset(2)};
      Stopwatch sw(hostTimes[1]);
      if (wb_add_las_file(g_ctx,h->getFileName().c_str(),h->getPointOffset(),h->numberPoints(),h->getPointFormat(),
                          h->getPointLength(),sc,of,h->getUnit())!=WB_OK)
      {
        die("Error reading file");
        break;
      }
      addInputRun(h,0,h->numberPoints());
      break;
    }
    case ACT_COUNT:
    {
      uint64_t c[256];
      if (g_classified && wb_count_classes(g_ctx,c)==WB_OK)
        for (int i=0;i<256;i++)
          if (c[i])
            classTotals[i]+=c[i];
      g_results.push_back(a);
      break;
    }
    default:
      break;
  }
}

void waitForQueueEmpty()
{
  flushPointBuffer();
  if (!g_files.empty())
    ensureBuilt();
}

namespace
{
struct RankBarrier
// the ranks meet here before the first collective: one that failed on its own (no device, unreadable file) makes
// everybody turn back instead of leaving the others inside NCCL
{
  mutex m;
  condition_variable cv;
  int world,arrived=0,failed=0;
  unsigned gen=0;
  explicit RankBarrier(int w):world(w) {}
  bool meet(bool ok)
  {
    unique_lock<mutex> lk(m);
    if (!ok)
      failed++;
    unsigned g=gen;
    if (++arrived==world)
    {
      arrived=0;
      gen++;
      cv.notify_all();
    }
    else
      cv.wait(lk,[&]{ return gen!=g; });
    return failed==0;
  }
};

bool shardable()
// every run is a whole file.  With at least as many files as GPUs the files are dealt out as x-strips (ascending x:
// wb_shard_run checks the order); with fewer — one big file — every worker reads every file and keeps its x-interval.
{
  if (g_nGpus<2 || g_files.empty())
    return false;
  for (auto &f:g_files)
    if (f.firstRec!=0 || f.count!=f.h->numberPoints())
      return false;
  return true;
}

bool classifySharded()
// One worker per GPU: each reads the files of its strip, takes part in wb_shard_run and hands back the class bytes
// of its own records; they go into the (built) store of the first device with wb_set_labels, so that counting,
// queries and the writers work as after a single-GPU classify.
{
  const bool windows=g_files.size()<(size_t)g_nGpus;
  const int W=windows?g_nGpus:(int)min<size_t>(g_nGpus,g_files.size());
  const auto t0=chrono::steady_clock::now();
  size_t total=0;
  for (auto &f:g_files)
    total+=f.count;
  // x-intervals for the window mode: equal-count quantiles of a sample of the records' x, moved onto boundaries of
  // the octree's finest cells (side/2^21) — two points with the same Morton key then always fall into the same
  // interval, so the order of equal keys (input order) is the single-GPU one on every rank
  vector<double> cuts;
  auto xOf=[](const LasHeader *h,size_t i)
  {
    int32_t X;
    memcpy(&X,h->records()+i*(size_t)h->getPointLength(),4);
    return (h->rawOffset(0)+h->rawScale(0)*(double)X)*h->getUnit();       // las.cpp:808, as wb_coord on the device
  };
  if (windows)
  {
    vector<double> sample;
    const size_t step=max<size_t>(1,total/2000000);
    for (auto &f:g_files)
      for (size_t i=0;i<f.count;i+=step)
        sample.push_back(xOf(f.h,i));
    sort(sample.begin(),sample.end());
    const double side=octRoot.getSide(),corner=octRoot.getCenter().getx()-0.5*side,cell=side/2097152.0;
    cuts.assign(W+1,0);
    cuts[0]=-INFINITY;
    cuts[W]=INFINITY;
    for (int r=1;r<W;r++)
    {
      double c=sample[sample.size()*(size_t)r/W];
      if (cell>0)
        c=corner+floor((c-corner)/cell+0.5)*cell;
      cuts[r]=c;
      if (!(cuts[r]>cuts[r-1]))
      {
        cerr<<"the points do not spread over "<<W<<" x-intervals\n";
        return false;
      }
    }
  }
  // contiguous groups of files, about total/W records each, none empty
  vector<size_t> firstFile(W+1,g_files.size());
  if (windows)
    for (int r=0;r<=W;r++)
      firstFile[r]=r<W?0:g_files.size();             // every worker walks all files
  else
  {
    size_t cum=0,i=0;
    for (int r=0;r<W;r++)
    {
      firstFile[r]=i;
      const size_t target=(size_t)((double)total*(r+1)/W);
      do
        cum+=g_files[i++].count;
      while (i<g_files.size()-(W-1-r) && cum+g_files[i].count/2<=target);
    }
    firstFile[W]=g_files.size();
  }
  const char *tr=getenv("WOLKEN_TRANSPORT");
  const bool local=tr && !strcmp(tr,"local");
  const char *d0=getenv("WOLKEN_DEVICE");
  const int base=d0?atoi(d0):0;
  int ndev=1;
  wb_device_count(&ndev);
  uint8_t id[WB_COMM_ID_BYTES];
  wb_local_group *grp=nullptr;
  if (local)
  {
    if (wb_local_group_create(W,&grp)!=WB_OK)
      return false;
  }
  else if (wb_comm_get_id(id)!=WB_OK)
  {
    cerr<<"NCCL is not available (libnccl.so.2; set WB_NCCL_LIB)\n";
    return false;
  }
  vector<uint8_t> labels(total);
  vector<vector<uint8_t>> rankLabels(windows?W:0);   // window mode: each worker's records are a subsequence of the input
  vector<string> errors(W);
  shardReport=ShardReport();
  shardReport.ranks.assign(W,wb_shard_stats());
  shardReport.files.assign(W,0);
  shardReport.device.assign(W,0);
  RankBarrier gate(W);
  auto work=[&](int r)
  {
    wb_ctx *c=nullptr;
    wb_comm *cm=nullptr;
    const int dev=local?(base+r)%max(1,ndev):base+r;
    bool ok=wb_create(dev,&c)==WB_OK;
    if (!ok)
      errors[r]="no CUDA device "+to_string(dev);
    size_t firstRecord=0;
    for (size_t i=0;i<firstFile[r];i++)
      firstRecord+=g_files[i].count;
    size_t nMine=0;
    if (ok)
    {
      wb_set_params(c,g_snakeTile,maxSlope,thickness,minHyperboloidSize);
      vector<size_t> mine;                           // the files this worker reads
      for (size_t i=windows?0:firstFile[r];ok && i<(windows?g_files.size():firstFile[r+1]);i++)
      {
        LasHeader *h=g_files[i].h;
        xyz a=h->minCorner(),b=h->maxCorner();
        double mn[3]={a.getx(),a.gety(),a.getz()},mx[3]={b.getx(),b.gety(),b.getz()};
        if (windows)
        {
          // the part of the file's box inside the interval: all workers' boxes together span the files' own boxes,
          // so the geometry (octree cube, tile lattice) is the single-GPU one
          mn[0]=max(mn[0],cuts[r]);
          mx[0]=min(mx[0],cuts[r+1]);
          if (!(mn[0]<=mx[0]))
            continue;                                // nothing of this file can lie in the interval
        }
        mine.push_back(i);
        ok=wb_add_extent(c,mn,mx)==WB_OK;
      }
      if (ok && windows)
        ok=wb_set_window(c,cuts[r],cuts[r+1])==WB_OK;
      if (ok && mine.empty())
      {
        ok=false;
        errors[r]="no input file reaches this worker's x-interval";
      }
      for (size_t k=0;ok && k<mine.size();k++)
      {
        const size_t i=mine[k];
        LasHeader *h=g_files[i].h;
        double sc[3]={h->rawScale(0),h->rawScale(1),h->rawScale(2)},of[3]={h->rawOffset(0),h->rawOffset(1),h->rawOffset(2)};
        ok=wb_add_las_file(c,h->getFileName().c_str(),h->getPointOffset(),h->numberPoints(),h->getPointFormat(),
                           h->getPointLength(),sc,of,h->getUnit())==WB_OK;
      }
      if (!ok && errors[r].empty())
        errors[r]=wb_last_error(c);
      if (ok && windows)
      {
        uint64_t kept=0;                             // records inside the interval = what the context holds now
        ok=wb_num_loaded(c,&kept)==WB_OK;
        nMine=kept;
        rankLabels[r].resize(nMine+1);
      }
    }
    if (gate.meet(ok))
    {
      // from here on every call is collective: a failure cannot be survived by the others
      int rc=local?wb_comm_init_local(c,grp,r,&cm):wb_comm_init(c,id,r,W,&cm);
      if (rc==WB_OK)
        rc=wb_shard_run(c,cm);
      if (rc==WB_OK)
        rc=wb_shard_get_labels(c,windows?rankLabels[r].data():labels.data()+firstRecord);
      if (rc!=WB_OK)
      {
        cerr<<"GPU "<<dev<<": "<<wb_last_error(c)<<endl;
        exit(4);
      }
      wb_shard_get_stats(c,&shardReport.ranks[r]);
      shardReport.files[r]=windows?g_files.size():firstFile[r+1]-firstFile[r];
      shardReport.device[r]=dev;
    }
    if (cm)
      wb_comm_destroy(cm);
    if (c)
      wb_destroy(c);
  };
  vector<thread> th;
  for (int r=0;r<W;r++)
    th.emplace_back(work,r);
  for (auto &t:th)
    t.join();
  if (grp)
    wb_local_group_destroy(grp);
  for (int r=0;r<W;r++)
    if (!errors[r].empty())
    {
      cerr<<"GPU worker "<<r<<": "<<errors[r]<<endl;
      return false;
    }
  if (windows)
  {
    // record i of the input belongs to the worker whose interval holds its x; each worker's class bytes come in the
    // order of its records, which is the input's.  Two passes over chunks of the input, in parallel.
    const int T=8;
    vector<size_t> fileFirst(g_files.size()+1,0);
    for (size_t i=0;i<g_files.size();i++)
      fileFirst[i+1]=fileFirst[i]+g_files[i].count;
    auto rankOf=[&](double x){ return (int)(upper_bound(cuts.begin()+1,cuts.end()-1,x)-(cuts.begin()+1)); };
    vector<vector<size_t>> cnt(T,vector<size_t>(W,0));
    auto span=[&](int t,size_t &a,size_t &b){ a=total*(size_t)t/T; b=total*(size_t)(t+1)/T; };
    auto walk=[&](int t,bool fill,vector<size_t> pos)
    {
      size_t a,b;
      span(t,a,b);
      size_t f=upper_bound(fileFirst.begin(),fileFirst.end(),a)-fileFirst.begin()-1;
      for (size_t i=a;i<b;i++)
      {
        while (i>=fileFirst[f+1])
          f++;
        const int r=rankOf(xOf(g_files[f].h,i-fileFirst[f]));
        if (fill)
          labels[i]=rankLabels[r][pos[r]++];
        else
          cnt[t][r]++;
      }
    };
    {
      vector<thread> th2;
      for (int t=0;t<T;t++)
        th2.emplace_back(walk,t,false,vector<size_t>());
      for (auto &t:th2)
        t.join();
    }
    vector<vector<size_t>> start(T,vector<size_t>(W,0));
    for (int r=0;r<W;r++)
    {
      size_t acc=0;
      for (int t=0;t<T;t++)
      {
        start[t][r]=acc;
        acc+=cnt[t][r];
      }
      if (acc+1!=rankLabels[r].size())
      {
        cerr<<"internal: worker "<<r<<" kept "<<rankLabels[r].size()-1<<" records, the host counts "<<acc<<endl;
        return false;
      }
    }
    {
      vector<thread> th2;
      for (int t=0;t<T;t++)
        th2.emplace_back(walk,t,true,start[t]);
      for (auto &t:th2)
        t.join();
    }
  }
  if (wb_set_labels(g_ctx,labels.data())!=WB_OK)
  {
    die("labels");
    return false;
  }
  shardReport.world=W;
  shardReport.seconds=chrono::duration<double>(chrono::steady_clock::now()-t0).count();
  return true;
}
} // namespace

void waitForThreads(int newStatus)
{
  g_command=newStatus;
  if (shardable() && (newStatus==TH_SCAN || newStatus==TH_POSTSCAN || newStatus==TH_SPLIT))
  {
    // several GPUs: the three tile phases are one collective run (scan and postscan results stay on the workers)
    ensureBuilt();
    if (newStatus==TH_SPLIT && !g_classified)
    {
      if (!classifySharded())
        exit(4);
      g_scanned=g_postscanned=g_classified=true;
      g_labels.clear();
    }
    return;
  }
  switch (newStatus)
  {
    case TH_SCAN:
      ensureBuilt();
      if (!g_scanned)
      {
        if (wb_scan(g_ctx)!=WB_OK) { die("scan"); exit(4); }
        g_scanned=true;
        tiles.invalidate();
      }
      break;
    case TH_POSTSCAN:
      waitForThreads(TH_SCAN);
      g_command=TH_POSTSCAN;
      if (!g_postscanned)
      {
        if (wb_postscan(g_ctx)!=WB_OK) { die("postscan"); exit(4); }
        g_postscanned=true;
        tiles.invalidate();
      }
      break;
    case TH_SPLIT:
      waitForThreads(TH_POSTSCAN);
      g_command=TH_SPLIT;
      if (!g_classified)
      {
        if (wb_classify(g_ctx)!=WB_OK) { die("classify"); exit(4); }
        g_classified=true;
        g_labels.clear();
      }
      break;
    default:
      break;
  }
}

void scanCylinder(Eisenstein) { waitForThreads(TH_SCAN); }
void postscanCylinder(Eisenstein) { waitForThreads(TH_POSTSCAN); }
void classifyCylinder(Eisenstein) { waitForThreads(TH_SPLIT); }

wb_ctx *wolkenContext() { ensureContext(); return g_ctx; }
const char *wolkenLastError() { return g_ctx?wb_last_error(g_ctx):""; }

vector<uint8_t> wolkenLabels()
{
  ensureLabels();
  return g_labels;
}

// ---------------------------------------------------------------- writer (cloudoutput.cpp:119-246, reduced)
static string className(int n)
{
  static const char *names[]={"raw","nonground","ground","lowveg","medveg","highveg","building","lownoise","",
                              "water","rail","road","overlap","wireguard","conductor","tower","insulator","bridge",
                              "highnoise","overhead","ignground","snow"};
  if (n>=0 && n<22 && names[n][0])
    return names[n];
  return "class"+to_string(n);
}

namespace
{
struct OutFile
{
  string name;
  FILE *f=nullptr;
  uint64_t n=0,byReturn[16]={0};
  int32_t mn[3]={INT32_MAX,INT32_MAX,INT32_MAX},mx[3]={INT32_MIN,INT32_MIN,INT32_MIN};
};
}

int writeClassified(const deque<LasHeader> &inputs,const OutputOptions &opt,vector<string> *written)
// One output (or one per class), optionally split every pointsPerFile points; file names
// name[-class][-k].las with k zero-padded as the reference does (cloudoutput.cpp:135-160).
// Records keep the inputs' format, scale and offset (all inputs must agree) with the class byte
// replaced; the header is rewritten with the new counts and bounding box.
{
  if (inputs.empty() || !g_classified)
    return -1;
  const LasHeader &h0=inputs[0];
  for (auto &h:inputs)
    if (h.getPointFormat()!=h0.getPointFormat() || h.getPointLength()!=h0.getPointLength() ||
        h.rawScale(0)!=h0.rawScale(0) || h.rawScale(1)!=h0.rawScale(1) || h.rawScale(2)!=h0.rawScale(2) ||
        h.rawOffset(0)!=h0.rawOffset(0) || h.rawOffset(1)!=h0.rawOffset(1) || h.rawOffset(2)!=h0.rawOffset(2))
    {
      cerr<<"inputs differ in format, scale or offset: merging them needs re-quantisation, which is not implemented\n";
      return -2;
    }
  ensureLabels();
  const int fmt=h0.getPointFormat(),len=h0.getPointLength();
  const unsigned hs=h0.headerLength();
  uint64_t grand=g_labels.size();
  int nDigits=0;
  if (opt.pointsPerFile)
  {
    uint64_t quot=(grand+opt.pointsPerFile-1)/opt.pointsPerFile;
    if (quot) quot--;
    if (!quot) quot++;
    while (quot) { quot/=10; nDigits++; }
  }
  map<int,vector<OutFile>> files;      // class (or -1 for all) -> sequence of files
  auto fileFor=[&](int cls)->OutFile &
  {
    vector<OutFile> &v=files[cls];
    if (v.empty() || (opt.pointsPerFile && v.back().n>=opt.pointsPerFile))
    {
      OutFile o;
      char num[32]="";
      if (opt.pointsPerFile)
        snprintf(num,sizeof(num),"-%0*zu",nDigits,v.size());
      o.name=opt.baseName+(cls>=0?"-"+className(cls):"")+num+".las";
      o.f=fopen(o.name.c_str(),"wb");
      if (o.f)
        fwrite(h0.headerBytes(),1,hs,o.f);
      v.push_back(o);
    }
    return v.back();
  };
  vector<uint8_t> rec(len);
  size_t idx=0;
  for (auto &h:inputs)
  {
    const uint8_t *r=h.records();
    for (size_t i=0;i<h.numberPoints();i++,idx++,r+=len)
    {
      uint8_t lab=g_labels[idx];
      OutFile &o=fileFor(opt.separateClasses?lab:-1);
      if (!o.f)
        return -3;
      memcpy(rec.data(),r,len);
      if (fmt<6)
        rec[15]=(uint8_t)((rec[15]&0xe0)|(lab&31));     // writePoint, las.cpp:848
      else
        rec[16]=lab;                                     // las.cpp:857
      fwrite(rec.data(),1,len,o.f);
      o.n++;
      int ret=fmt<6?(rec[14]&7):(rec[14]&15);
      if (ret>0 && ret<16)
        o.byReturn[ret]++;
      for (int k=0;k<3;k++)
      {
        int32_t v=rd<int32_t>(rec.data()+4*k);
        o.mn[k]=min(o.mn[k],v);
        o.mx[k]=max(o.mx[k],v);
      }
    }
  }
  const char *sysId=opt.separateClasses?"EXTRACTION":(inputs.size()>1?"MERGE":"MODIFICATION");   // cloudoutput.cpp:127-132
  for (auto &kv:files)
    for (auto &o:kv.second)
    {
      vector<uint8_t> hd(h0.headerBytes(),h0.headerBytes()+hs);
      memset(&hd[26],0,32);
      memcpy(&hd[26],sysId,strlen(sysId));
      memset(&hd[58],0,32);
      memcpy(&hd[58],"wolkenbase_b200",15);
      uint32_t off=hs,zero=0;
      memcpy(&hd[96],&off,4);
      memcpy(&hd[100],&zero,4);
      bool legacyOk=o.n<=4294967295ull && fmt<6;
      for (int i=0;i<6;i++)
      {
        uint32_t v=legacyOk?(uint32_t)(i?o.byReturn[i]:o.n):0;
        memcpy(&hd[107+4*i],&v,4);
      }
      for (int k=0;k<3;k++)
      {
        double mx=h0.rawOffset(k)+h0.rawScale(k)*o.mx[k],mn=h0.rawOffset(k)+h0.rawScale(k)*o.mn[k];
        memcpy(&hd[179+16*k],&mx,8);
        memcpy(&hd[187+16*k],&mn,8);
      }
      if (hs>=375)
      {
        uint64_t z=0,evlr=hs+o.n*len;
        memcpy(&hd[227],&z,8);
        memcpy(&hd[235],&evlr,8);
        memcpy(&hd[243],&zero,4);
        for (int i=0;i<16;i++)
        {
          uint64_t v=i?o.byReturn[i]:o.n;
          memcpy(&hd[247+8*i],&v,8);
        }
      }
      fseek(o.f,0,SEEK_SET);
      fwrite(hd.data(),1,hs,o.f);
      fclose(o.f);
      if (written)
        written->push_back(o.name);
    }
  return 0;
}

// ---------------------------------------------------------------- the reference's writer, restated
static const short kPointLengths[]={20,28,26,34,57,63,30,36,38,59,67};           // las.cpp:38
static const short kPointFeatures[]={0x0,0x1,0x2,0x3,0x9,0xb,0x101,0x103,0x107,0x109,0x10f};

int joinPointFormat(vector<int> formats)
{
  int all=0,ret=0;
  for (int f:formats)
    if (f>=0 && f<11)
      all|=kPointFeatures[f];
  for (int i=0;i<11;i++)
    if ((all|kPointFeatures[i])==kPointFeatures[i])
    {
      ret=i;
      break;
    }
  return ret;
}

static double pairwiseSum(const vector<double> &a)
// manysum.cpp:120-154: perfect trees over aligned power-of-two blocks, merged smallest block first
{
  double lv[32],sum=0;
  size_t n=a.size();
  for (size_t i=0;i<n;i++)
  {
    double v=a[i];
    size_t m=i;
    int l=0;
    while (m&1) { v=lv[l]+v; m>>=1; l++; }
    lv[l]=v;
  }
  for (int l=0;l<32;l++)
    if ((n>>l)&1)
      sum+=lv[l];
  return sum;
}

static void bubbleUp(vector<double> &v)
{
  for (size_t i=0;i+1<v.size();i++)
    if (v[i]>v[i+1])
      swap(v[i],v[i+1]);
}

static void bubbleDown(vector<double> &v)
{
  for (int i=(int)v.size()-2;i>=0;i--)
    if (v[i]>v[i+1])
      swap(v[i],v[i+1]);
}

static double manyGcd(vector<double> numbers,double toler)
// manygcd.cpp:41-76
{
  for (auto &x:numbers)
    x=fabs(x);
  if (numbers.empty())
    return 0;
  bubbleUp(numbers);
  bubbleDown(numbers);
  bubbleUp(numbers);
  while (numbers.size() && numbers.back()>numbers[0]+toler)
  {
    size_t sz=numbers.size();
    if (numbers[sz-1]-numbers[sz-2]>toler)
    {
      numbers[sz-1]-=numbers[sz-2];
      bubbleDown(numbers);
    }
    else
      numbers.resize(sz-1);
    bubbleUp(numbers);
  }
  return numbers.size()?pairwiseSum(numbers)/numbers.size():0;
}

static double combine1Scale(vector<double> &scales,vector<double> &offsets)
// las.cpp:906-925
{
  double toler=INFINITY,ret=0;
  for (double s:scales)
    if (fabs(s)<toler)
      toler=fabs(s);
  toler/=1e6;
  double scalegcd=manyGcd(scales,toler),offsetgcd=manyGcd(offsets,toler);
  if ((offsetgcd>10.5*scalegcd || offsetgcd==0) && fabs(offsetgcd/scalegcd-rint(offsetgcd/scalegcd))<1e-6)
    ret=scalegcd;
  for (int i=10;i>0;i--)
    for (int j=10;j>0;j--)
      if (fabs(scalegcd*i-offsetgcd*j)<toler)
        ret=scalegcd/j;
  return ret;
}

xyz combineScales(const deque<LasHeader> &headers)
{
  vector<double> xs,xo,ys,yo,zs,zo;
  for (auto &h:headers)
  {
    xyz s=h.getScale(),o=h.getOffset();
    xs.push_back(s.getx()); xo.push_back(o.getx());
    ys.push_back(s.gety()); yo.push_back(o.gety());
    zs.push_back(s.getz()); zo.push_back(o.getz());
  }
  return xyz(combine1Scale(xs,xo),combine1Scale(ys,yo),combine1Scale(zs,zo));
}

void LasHeader::openWrite(string fileName,int sysId)
{
  close();
  filename=fileName;
  out=fopen(fileName.c_str(),"wb+");
  switch (sysId)
  {
    case SI_MERGE: systemId="MERGE"; break;
    case SI_MODIFY: systemId="MODIFICATION"; break;
    case SI_EXTRACT: systemId="EXTRACTION"; break;
    case SI_TEST: systemId="TEST"; break;
    default: systemId="OTHER";
  }
  versionMajor=1;
  versionMinor=4;
  xScale=yScale=zScale=0;
  xOffset=yOffset=zOffset=NAN;
  pointFormat=pointLength=0;
  maxX=maxY=maxZ=-INFINITY;
  minX=minY=minZ=INFINITY;
  for (auto &v:nPoints)
    v=0;
}

void LasHeader::setVersion(int major,int minor)
{
  versionMajor=major;
  versionMinor=minor;
  headerSize=major==1?(minor<4?0xe3:0x177):0;
  writePos=pointOffset=headerSize;
}

void LasHeader::setPointFormat(int format)
{
  pointFormat=format;
  pointLength=(format>=0 && format<11)?kPointLengths[format]:0;
}

void LasHeader::setScale(xyz minCor,xyz maxCor,xyz scale)
// las.cpp:636-673
{
  auto one=[&](double mn,double mx,double sc,double &outScale,double &outOffset)
  {
    double minScale=(mx-mn)/4132485216.;
    outOffset=(mn+mx)/2/unit;
    if (sc>minScale && std::isfinite(sc))
    {
      outScale=sc/unit;
      outOffset=rint(outOffset/outScale)*outScale;
    }
    else
      outScale=minScale/unit;
  };
  one(minCor.getx(),maxCor.getx(),scale.getx(),xScale,xOffset);
  one(minCor.gety(),maxCor.gety(),scale.gety(),yScale,yOffset);
  one(minCor.getz(),maxCor.getz(),scale.getz(),zScale,zOffset);
}

static double binToDeg(int angle)
{
  return angle/2147483648.*360;
}

template <typename T> static void put(vector<uint8_t> &b,size_t o,T v)
{
  memcpy(&b[o],&v,sizeof(T));
}

void LasHeader::writePoint(const LasPoint &pnt)
// las.cpp:822-904
{
  if (!out)
    return;
  vector<uint8_t> r(pointLength,0);
  int xi=(int)lrint((pnt.location.getx()/unit-xOffset)/xScale);
  int yi=(int)lrint((pnt.location.gety()/unit-yOffset)/yScale);
  int zi=(int)lrint((pnt.location.getz()/unit-zOffset)/zScale);
  put<int32_t>(r,0,xi); put<int32_t>(r,4,yi); put<int32_t>(r,8,zi);
  double wx=xi*xScale+xOffset,wy=yi*yScale+yOffset,wz=zi*zScale+zOffset;
  if (wx>maxX) maxX=wx;
  if (wx<minX) minX=wx;
  if (wy>maxY) maxY=wy;
  if (wy<minY) minY=wy;
  if (wz>maxZ) maxZ=wz;
  if (wz<minZ) minZ=wz;
  put<uint16_t>(r,12,pnt.intensity);
  size_t o;
  if (pointFormat<6)
  {
    r[14]=(uint8_t)((pnt.returnNum&7)+((pnt.nReturns&7)<<3)+((pnt.scanDirection&1)<<6)+((pnt.edgeLine&1)<<7));
    r[15]=(uint8_t)((pnt.classification&31)+((pnt.classificationFlags&7)<<5));
    r[16]=(uint8_t)lrint(binToDeg(pnt.scanAngle));
    r[17]=(uint8_t)pnt.userData;
    put<uint16_t>(r,18,pnt.pointSource);
    o=20;
  }
  else
  {
    r[14]=(uint8_t)((pnt.returnNum&15)+((pnt.nReturns&15)<<4));
    r[15]=(uint8_t)((pnt.classificationFlags&15)+((pnt.scannerChannel&3)<<4)+((pnt.scanDirection&1)<<6)+((pnt.edgeLine&1)<<7));
    r[16]=(uint8_t)pnt.classification;
    r[17]=(uint8_t)pnt.userData;
    put<int16_t>(r,18,(int16_t)lrint(binToDeg(pnt.scanAngle)/0.006));
    put<uint16_t>(r,20,pnt.pointSource);
    o=22;
  }
  if ((1<<pointFormat)&0x7fa)
  {
    put<double>(r,o,pnt.gpsTime);
    o+=8;
  }
  if ((1<<pointFormat)&0x5ac)
  {
    put<uint16_t>(r,o,pnt.red); put<uint16_t>(r,o+2,pnt.green); put<uint16_t>(r,o+4,pnt.blue);
    o+=6;
  }
  if ((1<<pointFormat)&0x500)
    put<uint16_t>(r,o,pnt.nir);
  fseek(out,(long)writePos,SEEK_SET);
  fwrite(r.data(),1,r.size(),out);
  nPoints[0]++;
  if (pnt.returnNum>0 && pnt.returnNum<16)
    nPoints[pnt.returnNum]++;
  writePos+=pointLength;
}

int LasHeader::writeEncoded(wb_ctx *ctx,uint64_t arenaOff,size_t nBytes,const wb_file_stats &st)
// the effect of writePoint (las.cpp:822-904) for a run of records wb_encode has made on the device
{
  if (!out || !st.n_points[0])
    return 0;
  fflush(out);
  int rc=wb_write_encoded(ctx,fileno(out),writePos,arenaOff,nBytes);
  if (rc)
    return rc;
  writePos+=nBytes;
  for (int i=0;i<16;i++)
    nPoints[i]+=st.n_points[i];
  // wx=xi*scale+offset is monotone in xi: the extremes of the integers give the extremes of wx
  double lo[3]={st.imin[0]*xScale+xOffset,st.imin[1]*yScale+yOffset,st.imin[2]*zScale+zOffset};
  double hi[3]={st.imax[0]*xScale+xOffset,st.imax[1]*yScale+yOffset,st.imax[2]*zScale+zOffset};
  if (hi[0]>maxX) maxX=hi[0];
  if (lo[0]<minX) minX=lo[0];
  if (hi[1]>maxY) maxY=hi[1];
  if (lo[1]<minY) minY=lo[1];
  if (hi[2]>maxZ) maxZ=hi[2];
  if (lo[2]<minZ) minZ=lo[2];
  return 0;
}

void LasHeader::writeHeader()
// las.cpp:540-595.  Fields the reference never initialises on the write path (source id, global
// encoding, GUID, start of waveform data) are written as zeros.
{
  if (!out)
    return;
  vector<uint8_t> h(headerSize,0);
  memcpy(&h[0],"LASF",4);
  h[24]=(uint8_t)versionMajor;
  h[25]=(uint8_t)versionMinor;
  memcpy(&h[26],systemId.c_str(),min<size_t>(32,systemId.size()));
  const char *sw="wolkenbase_b200 (Wolkenbase 0.1.2~alpha)";
  memcpy(&h[58],sw,min<size_t>(32,strlen(sw)));
  time_t now=time(nullptr);
  tm *ptm=gmtime(&now);
  put<uint16_t>(h,90,(uint16_t)(ptm->tm_yday+1));
  put<uint16_t>(h,92,(uint16_t)(ptm->tm_year+1900));
  put<uint16_t>(h,94,(uint16_t)headerSize);
  put<uint32_t>(h,96,pointOffset);
  put<uint32_t>(h,100,0);
  h[104]=(uint8_t)pointFormat;
  put<uint16_t>(h,105,pointLength);
  bool legacyValid=true;
  for (int i=0;i<6;i++)
    if (nPoints[i]>4294967295ull)
      legacyValid=false;
  for (int i=6;i<16;i++)
    if (nPoints[i]>0)
      legacyValid=false;
  for (int i=0;i<6;i++)
    put<uint32_t>(h,107+4*i,legacyValid?(uint32_t)nPoints[i]:0);
  put<double>(h,131,xScale); put<double>(h,139,yScale); put<double>(h,147,zScale);
  put<double>(h,155,xOffset); put<double>(h,163,yOffset); put<double>(h,171,zOffset);
  put<double>(h,179,maxX); put<double>(h,187,minX); put<double>(h,195,maxY);
  put<double>(h,203,minY); put<double>(h,211,maxZ); put<double>(h,219,minZ);
  if (headerSize>0xe3)
  {
    put<uint64_t>(h,227,0);
    put<uint64_t>(h,235,(uint64_t)writePos);
    put<uint32_t>(h,243,0);
    for (int i=0;i<16;i++)
      put<uint64_t>(h,247+8*i,(uint64_t)nPoints[i]);
  }
  fseek(out,0,SEEK_SET);
  fwrite(h.data(),1,h.size(),out);
  fflush(out);
}

string CloudOutput::className(int n)
{
  return ::className(n);
}

static string nDecimal(long n,int dig)
{
  char buf[24]="";
  if (dig)
    snprintf(buf,sizeof(buf),"%0*ld",dig,n);
  return buf;
}

void CloudOutput::openFiles(string name,map<int,size_t> totals)
// cloudoutput.cpp:119-185
{
  int sysId=separateClasses?SI_EXTRACT:(nInputFiles>1?SI_MERGE:SI_MODIFY),nDigits=0;
  size_t quot;
  grandTotal=0;
  written.clear();
  for (auto &j:totals)
    grandTotal+=j.second;
  if (pointsPerFile)
  {
    quot=(grandTotal+pointsPerFile-1)/pointsPerFile;
    if (quot) quot--;
    if (!quot) quot++;
    while (quot) { quot/=10; nDigits++; }
  }
  auto openOne=[&](int cls,size_t i)
  {
    string fn=name+(cls>=0?"-"+className(cls):string())+(pointsPerFile?"-":"")+nDecimal((long)i,nDigits)+".las";
    deque<LasHeader> &d=headers[cls<0?0:cls];
    d.emplace_back();
    d.back().openWrite(fn,sysId);
    d.back().setUnit(unit);
    d.back().setScale(minCor,maxCor,scale);
    d.back().setVersion(1,4);
    d.back().setPointFormat(pointFormat);
    written.push_back(fn);
  };
  if (separateClasses)
    for (auto &j:totals)
    {
      quot=pointsPerFile?(j.second+pointsPerFile-1)/pointsPerFile:1;
      for (size_t i=0;i<quot;i++)
        openOne(j.first,i);
    }
  else
  {
    quot=pointsPerFile?(grandTotal+pointsPerFile-1)/pointsPerFile:1;
    for (size_t i=0;i<quot;i++)
      openOne(-1,i);
  }
}

void CloudOutput::writeFiles()
// cloudoutput.cpp:187-229: block by block, each class's points go to that class's least-full file
{
  int nextBlocks[256];
  for (size_t i=0;i<octStore.getNumBlocks();i++)
  {
    for (auto &k:headers)
    {
      long long mn=(long long)grandTotal;
      for (size_t j=0;j<k.second.size();j++)
        if ((long long)k.second[j].numberPoints()<mn)
        {
          nextBlocks[k.first]=(int)j;
          mn=(long long)k.second[j].numberPoints();
        }
    }
    for (auto &p:octStore.getAll((int64_t)i))
    {
      int cls=separateClasses?p.classification:0;
      auto it=headers.find(cls);
      if (it!=headers.end() && !it->second.empty())
        it->second[nextBlocks[cls]].writePoint(p);
    }
  }
}

void CloudOutput::writeCloudBlocks()
{
  int nextBlocks[256];
  for (int64_t i=0;i<getNumCloudBlocks();i++)
  {
    for (auto &k:headers)
    {
      long long mn=(long long)grandTotal;
      for (size_t j=0;j<k.second.size();j++)
        if ((long long)k.second[j].numberPoints()<mn)
        {
          nextBlocks[k.first]=(int)j;
          mn=(long long)k.second[j].numberPoints();
        }
    }
    for (auto &p:getCloudBlock(i))
    {
      int cls=separateClasses?p.classification:0;
      auto it=headers.find(cls);
      if (it!=headers.end() && !it->second.empty())
        it->second[nextBlocks[cls]].writePoint(p);
    }
  }
}

int CloudOutput::writeFilesDevice()
// writeFiles with the per-point work on the GPU: the sequential part (which file a bucket's points
// of one class go to, cloudoutput.cpp:197-206) stays here and needs only per-bucket class counts.
{
  ensureLeaves();
  vector<int> keys;                                  // class slots in map order
  vector<size_t> fileBase;                           // first global file index of each slot
  vector<LasHeader *> fileHdr;
  for (auto &k:headers)
  {
    keys.push_back(k.first);
    fileBase.push_back(fileHdr.size());
    for (auto &h:k.second)
      fileHdr.push_back(&h);
  }
  const size_t K=keys.size(),nl=g_leaves.size(),nf=fileHdr.size();
  if (!K || !nf || !nl)
    return 0;
  wb_out_spec spec;
  memset(&spec,0,sizeof(spec));
  spec.format=pointFormat;
  spec.rec_len=fileHdr[0]->getPointLength();
  spec.separate=separateClasses?1:0;
  spec.n_classes=separateClasses?(int)K:1;
  for (size_t k=0;k<K;k++)
    spec.classes[k]=(uint8_t)keys[k];
  for (int k=0;k<3;k++)
  {
    spec.scale[k]=fileHdr[0]->rawScale(k);
    spec.offset[k]=fileHdr[0]->rawOffset(k);
  }
  spec.unit=unit;
  vector<uint32_t> counts(nl*K);
  if (wb_leaf_class_counts(g_ctx,spec.classes,spec.n_classes,spec.separate,counts.data())!=WB_OK)
  {
    die("leaf class counts");
    return -1;
  }
  vector<uint64_t> filePts(nf,0),start(nl*K);
  vector<uint32_t> fileOf(nl*K);
  vector<int> next(K,0);
  for (size_t i=0;i<nl;i++)
  {
    for (size_t k=0;k<K;k++)
    {
      long long mn=(long long)grandTotal;
      size_t m=headers[keys[k]].size();
      for (size_t j=0;j<m;j++)
        if ((long long)filePts[fileBase[k]+j]<mn)
        {
          next[k]=(int)j;
          mn=(long long)filePts[fileBase[k]+j];
        }
    }
    for (size_t k=0;k<K;k++)
    {
      size_t f=fileBase[k]+next[k];
      fileOf[i*K+k]=(uint32_t)f;
      start[i*K+k]=filePts[f];
      filePts[f]+=counts[i*K+k];
    }
  }
  vector<uint64_t> fileOff(nf+1,0);
  for (size_t f=0;f<nf;f++)
    fileOff[f+1]=fileOff[f]+filePts[f]*(uint64_t)spec.rec_len;
  vector<uint64_t> dest(nl*K);
  for (size_t i=0;i<nl*K;i++)
    dest[i]=fileOff[fileOf[i]]+start[i]*(uint64_t)spec.rec_len;
  uint64_t total=fileOff[nf];
  vector<wb_file_stats> st(nf);
  // the records stay on the device; each file's span is streamed into it (pinned ring + pwrite threads)
  Stopwatch sw(hostTimes[3]);
  int rc=wb_encode(g_ctx,&spec,dest.data(),fileOf.data(),(uint32_t)nf,nullptr,total,st.data());
  if (rc!=WB_OK)
    die("encode");
  else
    for (size_t f=0;f<nf && rc==WB_OK;f++)
    {
      rc=fileHdr[f]->writeEncoded(g_ctx,fileOff[f],(size_t)(fileOff[f+1]-fileOff[f]),st[f]);
      if (rc!=WB_OK)
        die("write");
    }
  return rc==WB_OK?0:-1;
}

void CloudOutput::closeFiles()
{
  for (auto &k:headers)
  {
    for (auto &h:k.second)
    {
      h.writeHeader();
      h.close();
    }
    k.second.clear();
  }
  headers.clear();
}

int writeReferenceStyle(const deque<LasHeader> &inputs,const OutputOptions &opt,vector<string> *written)
// WolkenCanvas::writeFile (wolkencanvas.cpp:582-625) without the GUI: joined point format, combined
// scale, bounding box of the header corners, then CloudOutput open/write/close.
{
  if (inputs.empty() || !g_classified)
    return -1;
  vector<int> formats;
  vector<double> c;
  for (auto &h:inputs)
  {
    formats.push_back(h.getPointFormat());
    xyz a=h.minCorner(),b=h.maxCorner();
    c.push_back(a.getx()); c.push_back(a.gety()); c.push_back(a.getz());
    c.push_back(b.getx()); c.push_back(b.gety()); c.push_back(b.getz());
  }
  double box[6];
  wb_bound_rect(c.data(),(int)(c.size()/3),box);
  cloudOutput.pointFormat=joinPointFormat(formats);
  cloudOutput.minCor=xyz(box[0],box[1],box[2]);
  cloudOutput.maxCor=xyz(box[3],box[4],box[5]);
  cloudOutput.scale=combineScales(inputs);
  cloudOutput.nInputFiles=(int)inputs.size();
  cloudOutput.pointsPerFile=(int)opt.pointsPerFile;
  cloudOutput.separateClasses=opt.separateClasses;
  cloudOutput.unit=inputs[0].getUnit();
  map<int,size_t> totals=classTotals;
  if (totals.empty())
  {
    ensureLabels();
    for (uint8_t l:g_labels)
      totals[l]++;
  }
  if (!cloud.empty())
    totals[0]+=cloud.size();                         // getCloudBlock's points carry the default class
  cloudOutput.openFiles(opt.baseName,totals);
  int rc=0;
  if (keepRecordsOnDevice)
    rc=cloudOutput.writeFilesDevice();
  else
    cloudOutput.writeFiles();
  if (!rc)
    cloudOutput.writeCloudBlocks();
  cloudOutput.closeFiles();
  if (rc)
    return rc;
  if (written)
    *written=cloudOutput.written;
  return 0;
}
