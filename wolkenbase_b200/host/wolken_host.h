// wolken_host.h — the reference's C++ surface for the ground-extraction path, on top of the
// C ABI of libwolken_b200.so.  Same names, argument meaning and error behaviour as the
// reference headers it stands in for:
//   point.h (xy, xyz)            shape.h (Cube, Shape, Cylinder, Hyperboloid, Sphere, Paraboloid)
//   las.h (LasPoint, LasHeader)  octree.h (Octree octRoot, OctStore octStore)
//   eisenstein.h/flowsnake.h (Eisenstein, Flowsnake snake)   tile.h (Tile, tiles)
//   scan.h / classify.h (scanCylinder, postscanCylinder, classifyCylinder, tuning globals)
//   threads.h (startThreads, waitForThreads, enqueueAction, ... the phase protocol)
// What differs is WHERE the work runs: waitForThreads(TH_x) launches the phase on the GPU and
// returns when it is done, instead of flipping a command that worker threads poll.
#ifndef WOLKEN_HOST_H
#define WOLKEN_HOST_H
#include <cmath>
#include <cstdint>
#include <array>
#include <deque>
#include <fstream>
#include <map>
#include <string>
#include <vector>
#include "../../include/wolken_b200.h"

// ---------------------------------------------------------------- point.h
class xyz;
class xy
{
public:
  xy(double e=0,double n=0): x(e),y(n) {}
  xy(const xyz &p);
  double getx() const { return x; }
  double gety() const { return y; }
  double east() const { return x; }
  double north() const { return y; }
  double length() const { return hypot(x,y); }
  friend xy operator+(const xy &l,const xy &r) { return xy(l.x+r.x,l.y+r.y); }
  friend xy operator-(const xy &l,const xy &r) { return xy(l.x-r.x,l.y-r.y); }
  friend double dist(xy a,xy b) { return hypot(a.x-b.x,a.y-b.y); }
private:
  double x,y;
  friend class xyz;
};

class xyz
{
public:
  xyz(double e=0,double n=0,double h=0): x(e),y(n),z(h) {}
  xyz(xy en,double h): x(en.getx()),y(en.gety()),z(h) {}
  double getx() const { return x; }
  double gety() const { return y; }
  double getz() const { return z; }
  double east() const { return x; }
  double north() const { return y; }
  double elev() const { return z; }
  bool isnan() const { return std::isnan(x) || std::isnan(y) || std::isnan(z); }
  bool isfinite() const { return std::isfinite(x) && std::isfinite(y) && std::isfinite(z); }
  friend xyz operator+(const xyz &l,const xyz &r) { return xyz(l.x+r.x,l.y+r.y,l.z+r.z); }
  friend xyz operator-(const xyz &l,const xyz &r) { return xyz(l.x-r.x,l.y-r.y,l.z-r.z); }
  friend xyz operator*(const xyz &l,double r) { return xyz(l.x*r,l.y*r,l.z*r); }
  friend bool operator==(const xyz &l,const xyz &r) { return l.x==r.x && l.y==r.y && l.z==r.z; }
private:
  double x,y,z;
  friend class xy;
};
inline xy::xy(const xyz &p): x(p.x),y(p.y) {}
inline double sqr(double v) { return v*v; }

// ---------------------------------------------------------------- shape.h (shape.cpp:26-274)
class Cube
{
public:
  Cube(): side(0) {}
  Cube(xyz c,double s): center(c),side(s) {}
  bool in(xyz pnt) const;
  xyz getCenter() const { return center; }
  double getSide() const { return side; }
  xyz corner(int n) const;
private:
  xyz center;
  double side;
};

class Shape
{
public:
  virtual ~Shape() {}
  virtual bool in(xyz pnt) const=0;
  virtual bool in(Cube &cube) const;               // all eight corners (convex shapes)
  virtual xyz closestPoint(Cube cube) const=0;
  virtual bool intersect(Cube cube) const { return in(closestPoint(cube)); }
  virtual bool abi(wb_shape &) const { return false; } // constructor arguments for the device queries; false = CPU only
};

class Paraboloid: public Shape
{
public:
  Paraboloid(): radiusCurvature(0) {}
  Paraboloid(xyz v,double r): vertex(v),radiusCurvature(r) {}
  bool in(xyz pnt) const override;
  xyz closestPoint(Cube cube) const override;
  using Shape::in;
  bool abi(wb_shape &s) const override;
private:
  xyz vertex;
  double radiusCurvature;
};

class Hyperboloid: public Shape
{
public:
  Hyperboloid(): por2(0),slope(1),r0(0) {}
  Hyperboloid(xyz v,double r,double s);
  bool in(xyz pnt) const override;
  xyz closestPoint(Cube cube) const override;
  using Shape::in;
  bool abi(wb_shape &s) const override;
private:
  xyz center,vertex0;
  double por2,slope,r0;
};

class Sphere: public Shape
{
public:
  Sphere(): radius(0) {}
  Sphere(xyz c,double r): center(c),radius(r) {}
  bool in(xyz pnt) const override;
  xyz closestPoint(Cube cube) const override;
  using Shape::in;
  bool abi(wb_shape &s) const override;
private:
  xyz center;
  double radius;
};

class Cylinder: public Shape
{
public:
  Cylinder(): radius(0) {}
  Cylinder(xy c,double r): center(c),radius(r) {}
  double getRadius() const { return radius; }
  xy getCenter() const { return center; }
  bool in(xyz pnt) const override;
  xyz closestPoint(Cube cube) const override;
  using Shape::in;
  bool abi(wb_shape &s) const override;
private:
  xy center;
  double radius;
};

class Column: public Shape                          // shape.cpp:240-274
{
public:
  Column(): side(0) {}
  Column(xy c,double s): center(c),side(s) {}
  bool in(xyz pnt) const override;
  xyz closestPoint(Cube cube) const override;
  using Shape::in;
  bool abi(wb_shape &s) const override;
private:
  xy center;
  double side;
};

// ---------------------------------------------------------------- las.h
class LasPoint
{
public:
  xyz location;
  unsigned short intensity,returnNum,nReturns;
  bool scanDirection,edgeLine;
  unsigned short classification,classificationFlags,scannerChannel,userData,pointSource;
  int scanAngle;
  double gpsTime;
  unsigned short nir,red,green,blue;
  LasPoint();
  bool isEmpty() const { return location.isnan(); }
};

class LasHeader
// Read side of las.cpp:299-428, 735-820 over a memory-mapped file (the records are handed to
// the GPU as one span); write side: header + patched records (las.cpp:540-595).
{
public:
  LasHeader();
  ~LasHeader();
  LasHeader(const LasHeader &o);                    // the reference copies headers into deques; only closed/unopened ones
  LasHeader &operator=(const LasHeader &)=delete;
  void openRead(std::string fileName);
  bool isValid() const;
  bool isZipped() const { return zipFlag; }
  void close();
  void setUnit(double u) { unit=u; }
  double getUnit() const { return unit; }
  std::string getFileName() const { return filename; }
  size_t numberPoints(int r=0) const { return nPoints[r]; }
  int getVersion() const { return (versionMajor<<8)+versionMinor; }
  int getPointFormat() const { return pointFormat; }
  int getPointLength() const { return pointLength; }
  size_t getPointOffset() const { return pointOffset; }
  xyz getScale() const { return xyz(xScale*unit,yScale*unit,zScale*unit); }
  xyz getOffset() const { return xyz(xOffset*unit,yOffset*unit,zOffset*unit); }
  xyz minCorner() const { return xyz(minX*unit,minY*unit,minZ*unit); }
  xyz maxCorner() const { return xyz(maxX*unit,maxY*unit,maxZ*unit); }
  // write side (las.cpp:456-514, 540-595, 613-673, 822-904)
  void openWrite(std::string fileName,int sysId);
  void setVersion(int major,int minor);
  void setPointFormat(int format);
  void setScale(xyz minCor,xyz maxCor,xyz scale);
  void writePoint(const LasPoint &pnt);
  void writeHeader();
  int writeEncoded(wb_ctx *ctx,uint64_t arenaOff,size_t nBytes,const wb_file_stats &st);   // records made by wb_encode
  LasPoint readPoint(size_t num);                   // throws int -1 past the end, like the reference
  const uint8_t *records() const { return map?map+pointOffset:nullptr; }
  const uint8_t *headerBytes() const { return map; }
  unsigned headerLength() const { return headerSize; }
  double rawScale(int k) const { return k==0?xScale:(k==1?yScale:zScale); }
  double rawOffset(int k) const { return k==0?xOffset:(k==1?yOffset:zOffset); }
private:
  std::string filename;
  uint8_t *map;
  size_t mapLen;
  int versionMajor,versionMinor;
  unsigned headerSize,pointOffset;
  unsigned short pointFormat,pointLength;
  double xScale,yScale,zScale,xOffset,yOffset,zOffset,maxX,minX,maxY,minY,maxZ,minZ,unit;
  size_t nPoints[16];
  bool zipFlag;
  FILE *out;
  size_t writePos;
  std::string systemId;
};

#define SI_MERGE 0
#define SI_MODIFY 1
#define SI_EXTRACT 2
#define SI_TEST 3
int joinPointFormat(std::vector<int> formats);      // las.cpp:103-119
xyz combineScales(const std::deque<LasHeader> &headers);   // las.cpp:906-944

class CloudOutput                                    // cloudoutput.h:30-48, without Qt
{
public:
  xyz minCor,maxCor,scale;
  int nInputFiles=0;
  size_t grandTotal=0;
  int pointsPerFile=0;                               // 0 means no limit
  int pointFormat=0;
  bool separateClasses=true,writeLaz=false;
  double unit=1;
  std::string className(int n);
  void openFiles(std::string name,std::map<int,size_t> classTotals);
  void writeFiles();
  int writeFilesDevice();                           // the same files, records made by wb_encode
  void writeCloudBlocks();                          // cloudoutput.cpp:213-229: the points of `cloud` after the store's
  void closeFiles();
  std::vector<std::string> written;
private:
  std::map<int,std::deque<LasHeader> > headers;
};
extern CloudOutput cloudOutput;

// ---------------------------------------------------------------- eisenstein.h / flowsnake.h / tile.h
class Eisenstein
{
public:
  Eisenstein(int xa=0,int ya=0): x(xa),y(ya) {}
  int getx() const { return x; }
  int gety() const { return y; }
  friend bool operator<(const Eisenstein &a,const Eisenstein &b) { return a.y!=b.y?a.y<b.y:a.x<b.x; }
private:
  int x,y;
};

Eisenstein toFlowsnake(int n);                      // flowsnake.cpp:92-136

class Flowsnake
{
public:
  void setSize(Cube cube,double desiredSpacing);
  void restart() { counter=startnum; }
  Eisenstein next();                                // INT_MIN,INT_MIN when exhausted
  Cylinder cyl(Eisenstein e);
  double progress() { return 1; }                   // every phase is complete when its call returns
  double getSpacing() const { return spacing; }
  int first() const { return startnum; }
  int last() const { return stopnum; }
private:
  xy center;
  double spacing=0;
  int startnum=0,counter=0,stopnum=-1;
};

struct Tile                                          // tile.h:27-35
{
  int nPoints,nGround;
  short roofFlags,treeFlags;
  double density,hyperboloidSize,height;
};

void refreshTiles();
class TileTable                                      // stands in for harray<Tile> tiles
{
public:
  Tile &operator[](Eisenstein e);                   // zero tile if absent, like harray
  int count(Eisenstein e) { sync(); return byAddr.count(e); }
  void clear() { byAddr.clear(); stale=false; }
  size_t size() { sync(); return byAddr.size(); }
  void invalidate() { stale=true; }                 // the device table changed: refill on next access
private:
  void sync();
  bool stale=false;
  std::map<Eisenstein,Tile> byAddr;
  friend void refreshTiles();
};

// ---------------------------------------------------------------- octree.h
#define RECORDS WB_RECORDS

class Octree
{
public:
  void sizeFit(std::vector<xyz> pnts);              // octree.cpp:268-310
  xyz getCenter() const { return center; }
  double getSide() const { return side; }
  Cube cube() const { return Cube(center,side); }
  int64_t findBlock(xyz pnt);                       // leaf (= block) index in dump order, -1 if none
  Cube findCube(xyz pnt);
  std::vector<int64_t> findBlocks(const Shape &sh); // octree.cpp:234-251, same order
  void clear();
private:
  xyz center;
  double side=0;
  friend class OctStore;
};

class OctStore
{
public:
  size_t getNumBlocks();
  std::vector<LasPoint> getAll(int64_t block);
  std::vector<LasPoint> pointsIn(const Shape &sh,bool sorted=false);   // octree.cpp:1214-1242
  uint64_t countPointsIn(const Shape &sh);
  std::array<double,2> hiLoPointsIn(const Shape &sh);
  std::map<int,size_t> countClasses(int64_t block);
  void dump(std::ofstream &file);                   // octree.cpp:888-891
  uint64_t countPoints();
  void clear();
  void disown() {}
  void setIgnoreDupes(bool) {}
  void shrink() {}
};

// ---------------------------------------------------------------- cloud.h
// Points that came from a non-LAS source (the reference's PLY/XYZ readers, ply.cpp:54, fileio.cpp:81-94) wait here,
// outside the octree; the writer appends them, unclassified, after the store's blocks (cloudoutput.cpp:213-229).
extern std::vector<xyz> cloud;
int64_t getNumCloudBlocks();                        // cloud.cpp:29-32
std::vector<LasPoint> getCloudBlock(int64_t n);     // cloud.cpp:34-45

// ---------------------------------------------------------------- testpattern.h
// Test data carries its point number as GPS time.  censusPoints() walks the store and prints, as the reference does
// after every write (threads.cpp:613), "Duplicate point" if a number occurs twice, "Max point N", and the missing
// numbers below N; returns the number of missing points (-1: not test data).  With the records on the device
// (keepRecordsOnDevice) the walk is wb_census; otherwise block by block through getAll.
int censusPoints(std::vector<LasPoint> points);     // testpattern.cpp:56-82
long long censusPoints(std::ostream *report=nullptr);   // testpattern.cpp:84-123 (report: default std::cout)

extern Octree octRoot;
extern OctStore octStore;
extern Flowsnake snake;
extern TileTable tiles;
extern std::map<int,size_t> classTotals;
extern double minHyperboloidSize,maxSlope,thickness;  // scan.h:25
extern double tileSize;                               // the GUI's setting, mainwindow.cpp:398-411
extern double hostTimes[4];        // seconds spent in wb_create, wb_add_las_file, wb_build, wb_encode+write
extern bool hostQueries;           // true: OctStore queries run on the CPU mirror (the checker) instead of the GPU
extern bool keepRecordsOnDevice;   // set before reading: the raw records stay in device memory and ACT_WRITE's
                                   // records are made there (wb_encode) instead of by LasHeader::writePoint

// ---------------------------------------------------------------- threads.h
#define TH_WAIT 1
#define TH_READ 2
#define TH_SCAN 3
#define TH_POSTSCAN 4
#define TH_SPLIT 5
#define TH_PAUSE 6
#define TH_STOP 7
#define TH_ASLEEP 256
#define ACT_READ 1
#define ACT_COUNT 2
#define ACT_WRITE 3
#define ACT_LOAD 4                                  // accepted, no-op: there is no block file to load (threads.cpp:591-628)

struct ThreadAction
{
  int opcode=0;
  int param0=0;
  double param1=0,param2=0;
  LasHeader *hdr=nullptr;                           // borrowed; must outlive the read
  std::string filename;
  int flags=0,result=0;
};

void startThreads(int n);                           // n is accepted and ignored: one GPU context
void joinThreads();
void enqueueAction(ThreadAction a);                 // ACT_READ runs at once; ACT_COUNT fills classTotals
ThreadAction dequeueResult();
bool actionQueueEmpty();
bool resultQueueEmpty();
Eisenstein dequeueTileDone();                       // always (INT_MIN,INT_MIN): phases complete as a whole, no per-tile repaint queue
bool tileDoneQueueEmpty();
// The reference's wolkencli feeds the pool itself: readPoint, then embufferPoint (wolkencli.cpp:104-108).  Here the
// records such points came from wait as runs of consecutive records and reach the GPU in waitForQueueEmpty (or any
// phase change); a point that is not the one the last readPoint on this thread returned is refused with a message.
void embufferPoint(LasPoint point,bool fromFile);
void embufferPoints(std::vector<LasPoint> points,int thread);   // split re-insertion (octree.cpp:1332): never needed here
LasPoint debufferPoint(int thread);                 // always the empty point
bool pointBufferEmpty();
size_t pointBufferSize();                           // embuffered points not yet on the device
void sleepDead(int thread);
int thisThread();                                   // -1 = the calling (main) thread, threads.cpp:414-417
size_t duplicatePoints();                           // alreadyInOctree.size() (octree.cpp:620-662): records lost to an identical XYZ
void setThreadCommand(int newStatus);
int getThreadCommand();
int getThreadStatus();                              // (command<<20)|command: "all threads in the commanded state"
void waitForThreads(int newStatus);                 // runs the phase: TH_SCAN build+scan, TH_POSTSCAN, TH_SPLIT classify
void waitForQueueEmpty();
int nThreads();
double busyFraction();
void initTiles();

void scanCylinder(Eisenstein cylAddress);           // scan.h:27 — first call runs the whole phase
void postscanCylinder(Eisenstein cylAddress);
void classifyCylinder(Eisenstein cylAddress);       // classify.h:24
void fillTanTables();                               // uploaded by wb_create; kept for source compatibility

// ---------------------------------------------------------------- the CLI's additions
// startThreads(n) with n > 1 spreads scan, postscan and classify over n GPUs when the inputs are at least two whole
// files in ascending x (x-strips): what each worker did in the last such run.
struct ShardReport
{
  int world=0;
  double seconds=0;                                 // wall time of the whole collective run, file reads included
  std::vector<wb_shard_stats> ranks;
  std::vector<size_t> files;                        // input files per worker
  std::vector<int> device;
};
extern ShardReport shardReport;
wb_ctx *wolkenContext();
const char *wolkenLastError();
std::vector<uint8_t> wolkenLabels();                // class byte per input record (files in read order)
struct OutputOptions
{
  std::string baseName;
  bool separateClasses=true;                        // mainwindow.cpp:398-411 defaults
  size_t pointsPerFile=0;
};
// Lossless writer: the inputs' own records with the class byte replaced (no re-quantisation).
int writeClassified(const std::deque<LasHeader> &inputs,const OutputOptions &opt,std::vector<std::string> *written);
// The reference's writer (CloudOutput: LAS 1.4, joined point format, combined scale, re-quantised XYZ).
int writeReferenceStyle(const std::deque<LasHeader> &inputs,const OutputOptions &opt,std::vector<std::string> *written);
#endif
