/* ref_driver.cpp — headless driver for the UNMODIFIED reference (test infrastructure only).
 *
 * This file is ours; it contains no reference source.  It is linked against objects
 * compiled straight from /root/reference (see oracle/Makefile) and replays the phase
 * sequence the reference's GUI performs (wolkencanvas.cpp:502-580, tick 341-354):
 *   read+build -> [canonical re-insertion] -> scan -> postscan -> classify -> stop
 * and then writes the three parity artefacts:
 *   <out>.dump    octStore.dump()                     (octree.cpp:888-891)
 *   <out>.tiles   per non-empty tile: n, ex, ey, nPoints, treeFlags, density,
 *                 hyperboloidSize, height                (tile.h:27-35)
 *   <out>.labels  one class byte per point, indexed by lrint(gpsTime)
 * plus one JSON line with per-phase seconds on stdout.
 *
 * Canonical re-insertion: after the build, all points are pulled out,
 * sorted by (21-level Morton key from the reference's own ">=center" descent, gpsTime)
 * and pushed back unshuffled, so that OctStore::pointsIn returns points in a DEFINED
 * order (the scan phase is order dependent, scan.cpp:78-82).  Only meaningful with 1 thread.
 *
 * usage: ref_driver [-t threads] [-c] [-T tileSize] [-S maxSlope] [-K thickness]
 *                   [-M minHyperboloidSize] [-w lasprefix [-s 0|1] [-p pointsPerFile]]
 *                   -o outprefix in1.las [in2.las ...]
 * -w writes the classified cloud with the reference's own LasHeader write path
 * (las.cpp:456-514, 540-595, 636-673, 822-904), driven the way CloudOutput does it
 * (cloudoutput.cpp:119-246; that class itself needs Qt, so its three loops are replayed here).
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <climits>
#include <deque>
#include <map>
#include <vector>
#include <string>
#include <iostream>
#include <fstream>
#include <algorithm>
#include <chrono>
#include <thread>
#include <unistd.h>
#include "las.h"
#include "octree.h"
#include "threads.h"
#include "angle.h"
#include "scan.h"
#include "tile.h"
#include "flowsnake.h"
#include "relprime.h"
#include "freeram.h"
#include "boundrect.h"

using namespace std;
namespace cr=std::chrono;

static const double kSquareSides[12]= // flowsnake.cpp:30-44 is file-static: restated constants
{
  0.6583539906808145,1.8501627472990723,4.286014912881196,12.6716160058597,
  32.85016274729906,79.01914481623778,243.0343734125204,592.8501627472989,
  1510.850162747299,4399.956311577825,10731.850162747294,29198.8501627473
};

static double now()
{
  return cr::duration<double>(cr::steady_clock::now().time_since_epoch()).count();
}

static uint64_t mortonKey(xyz p,xyz c,double side)
/* 21-level key, 3 bits per level, child index i=z*4+y*2+x with bit = coord>=center
 * (octree.cpp:204-207), centre of child = centre +- side/4 (octree.cpp:335-337). */
{
  uint64_t key=0;
  double cx=c.getx(),cy=c.gety(),cz=c.getz(),q=side/4;
  for (int l=0;l<21;l++)
  {
    int xb=p.getx()>=cx,yb=p.gety()>=cy,zb=p.getz()>=cz;
    key=(key<<3)|(zb*4+yb*2+xb);
    cx+=(2*xb-1)*q;
    cy+=(2*yb-1)*q;
    cz+=(2*zb-1)*q;
    q/=2;
  }
  return key;
}

struct Keyed
{
  uint64_t key;
  double t;
  LasPoint p;
};

static void waitProgress()
{
  while (snake.progress()<1)
    this_thread::sleep_for(chrono::milliseconds(2));
  this_thread::sleep_for(chrono::milliseconds(60)); // let the last tile (n==stopnum) be taken
}

int main(int argc,char **argv)
{
  int nthreads=1,opt;
  bool canonical=false,separateClasses=true;
  double tileSize=1;
  long pointsPerFile=0;
  string out="ref",lasOut;
  maxSlope=1;
  thickness=0;
  minHyperboloidSize=0.1;
  while ((opt=getopt(argc,argv,"t:cT:S:K:M:o:w:s:p:"))!=-1)
    switch (opt)
    {
      case 't': nthreads=atoi(optarg); break;
      case 'c': canonical=true; break;
      case 'T': tileSize=atof(optarg); break;
      case 'S': maxSlope=atof(optarg); break;
      case 'K': thickness=atof(optarg); break;
      case 'M': minHyperboloidSize=atof(optarg); break;
      case 'o': out=optarg; break;
      case 'w': lasOut=optarg; break;
      case 's': separateClasses=atoi(optarg)!=0; break;
      case 'p': pointsPerFile=atol(optarg); break;
      default: return 2;
    }
  if (nthreads<1)
    nthreads=thread::hardware_concurrency();
  if (optind>=argc)
  {
    fprintf(stderr,"no input files\n");
    return 2;
  }
  deque<LasHeader> headers;
  size_t total=0;
  for (int i=optind;i<argc;i++)
  {
    headers.emplace_back();
    headers.back().openRead(argv[i]);
    if (!headers.back().isValid())
    {
      fprintf(stderr,"%s: not a valid LAS file\n",argv[i]);
      return 1;
    }
    total+=headers.back().numberPoints();
  }
  double t0=now();
  fillTanTables();
  lowRam=freeRam()/7;
  octStore.open(out+".store.oct",nthreads+relprime(nthreads));
  octStore.resize(8*nthreads+1);
  startThreads(nthreads);
  waitForThreads(TH_READ);
  vector<xyz> limits;
  BoundRect br;
  multimap<int64_t,LasHeader *> sorter;
  for (size_t i=0;i<headers.size();i++)
  {
    limits.push_back(headers[i].minCorner());
    limits.push_back(headers[i].maxCorner());
    br.include(headers[i].minCorner());
    br.include(headers[i].maxCorner());
    sorter.insert(pair<int64_t,LasHeader *>(-(int64_t)headers[i].numberPoints(),&headers[i]));
  }
  octRoot.sizeFit(limits);
  double side=br.right()-br.left();
  if (br.top()-br.bottom()>side)
    side=br.top()-br.bottom();
  if (br.high()-br.low()>side)
    side=br.high()-br.low();
  Cube cube(xyz((br.right()+br.left())/2,(br.top()+br.bottom())/2,(br.high()+br.low())/2),side);
  snake.setSize(cube,tileSize);
  initTiles();
  for (auto j=sorter.begin();j!=sorter.end();++j)
  {
    ThreadAction ta;
    ta.hdr=j->second;
    ta.opcode=ACT_READ;
    enqueueAction(ta);
  }
  this_thread::sleep_for(chrono::milliseconds(50));
  waitForQueueEmpty();
  double tRead=now();
  if (canonical)
  {
    if (nthreads!=1)
      fprintf(stderr,"warning: canonical order is only defined for 1 thread\n");
    waitForThreads(TH_WAIT);
    vector<Keyed> all;
    all.reserve(total);
    xyz ctr=octRoot.getCenter();
    double rside=octRoot.getSide();
    for (int64_t b=0;b<(int64_t)octStore.getNumBlocks();b++)
    {
      vector<LasPoint> blk=octStore.getAll(b);
      octStore.disown();
      for (size_t k=0;k<blk.size();k++)
      {
	Keyed kd;
	kd.key=mortonKey(blk[k].location,ctr,rside);
	kd.t=blk[k].gpsTime;
	kd.p=blk[k];
	all.push_back(kd);
      }
    }
    sort(all.begin(),all.end(),[](const Keyed &a,const Keyed &b)
	 {return a.key!=b.key?a.key<b.key:a.t<b.t;});
    octStore.clearBlocks();
    octStore.disown();
    vector<LasPoint> rev;
    rev.reserve(all.size());
    for (size_t k=all.size();k-->0;)
      rev.push_back(all[k].p);
    all.clear();
    all.shrink_to_fit();
    embufferPoints(rev,0);
    rev.clear();
    rev.shrink_to_fit();
    waitForThreads(TH_READ);
    this_thread::sleep_for(chrono::milliseconds(50));
    waitForQueueEmpty();
  }
  double tCanon=now();
  waitForThreads(TH_SCAN);
  octStore.shrink();
  waitProgress();
  double tScan=now();
  waitForThreads(TH_POSTSCAN);
  snake.restart();
  waitProgress();
  double tPost=now();
  waitForThreads(TH_SPLIT);
  octStore.setIgnoreDupes(true);
  snake.restart();
  waitProgress();
  waitForThreads(TH_PAUSE);
  double tClass=now();
  waitForThreads(TH_STOP);
  joinThreads();
  // ---- artefacts -------------------------------------------------------
  {
    ofstream dumpFile(out+".dump");
    octStore.dump(dumpFile);
  }
  vector<unsigned char> labels(total,255);
  size_t nStored=0,nBadTime=0;
  for (int64_t b=0;b<(int64_t)octStore.getNumBlocks();b++)
  {
    vector<LasPoint> blk=octStore.getAll(b);
    octStore.disown();
    for (size_t k=0;k<blk.size();k++)
    {
      long long t=llrint(blk[k].gpsTime);
      nStored++;
      if (t>=0 && (size_t)t<total)
	labels[t]=blk[k].classification;
      else
	nBadTime++;
    }
  }
  {
    ofstream lf(out+".labels",ios::binary);
    lf.write((const char *)labels.data(),labels.size());
  }
  if (lasOut.size())
  {
    // WolkenCanvas::writeFile (wolkencanvas.cpp:582-625) + CloudOutput::openFiles/writeFiles/closeFiles
    static const char *names[]={"raw","nonground","ground"};
    vector<int> formats;
    for (size_t i=0;i<headers.size();i++)
      formats.push_back(headers[i].getPointFormat());
    int pointFormat=joinPointFormat(formats);
    xyz minCor(br.left(),br.bottom(),br.low()),maxCor(br.right(),br.top(),br.high());
    xyz scale=combineScales(headers);
    map<int,size_t> totals;
    for (int64_t b=0;b<(int64_t)octStore.getNumBlocks();b++)
    {
      map<int,size_t> c=octStore.countClasses(b);
      octStore.disown();
      for (auto &j:c)
        totals[j.first]+=j.second;
    }
    size_t grandTotal=0,quot;
    int nDigits=0;
    for (auto &j:totals)
      grandTotal+=j.second;
    int sysId=separateClasses?SI_EXTRACT:(headers.size()>1?SI_MERGE:SI_MODIFY);
    if (pointsPerFile)
    {
      quot=(grandTotal+pointsPerFile-1)/pointsPerFile;
      if (quot) quot--;
      if (!quot) quot++;
      while (quot) { quot/=10; nDigits++; }
    }
    map<int,deque<LasHeader> > outs;
    auto openOne=[&](int cls,size_t i)
    {
      char num[32]="";
      if (nDigits)
        snprintf(num,sizeof(num),"%0*zu",nDigits,i);
      string fn=lasOut+(cls>=0?string("-")+(cls<3?names[cls]:"other"):string())+(pointsPerFile?"-":"")+num+".las";
      deque<LasHeader> &d=outs[cls<0?0:cls];
      d.push_back(LasHeader());
      d.back().openWrite(fn,sysId);
      d.back().setUnit(1);
      d.back().setScale(minCor,maxCor,scale);
      d.back().setVersion(1,4);
      d.back().setPointFormat(pointFormat);
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
    int nextBlocks[256];
    for (int64_t b=0;b<(int64_t)octStore.getNumBlocks();b++)
    {
      for (auto &k:outs)
      {
        long long mn=grandTotal;
        for (size_t j=0;j<k.second.size();j++)
          if ((long long)k.second[j].numberPoints()<mn)
          {
            nextBlocks[k.first]=j;
            mn=k.second[j].numberPoints();
          }
      }
      vector<LasPoint> blk=octStore.getAll(b);
      octStore.disown();
      for (size_t j=0;j<blk.size();j++)
      {
        int cls=separateClasses?blk[j].classification:0;
        outs[cls][nextBlocks[cls]].writePoint(blk[j]);
      }
    }
    for (auto &k:outs)
      for (size_t i=0;i<k.second.size();i++)
      {
        k.second[i].writeHeader();
        k.second[i].close();
      }
  }
  int sizeIndex=-1;
  for (int i=0;i<12;i++)
    if (cube.getSide()/kSquareSides[i]==snake.getSpacing())
      sizeIndex=i;
  size_t nTiles=0;
  if (sizeIndex>=0)
  {
    ofstream tf(out+".tiles",ios::binary);
    double hdr[4]={snake.getSpacing(),cube.getCenter().getx(),cube.getCenter().gety(),(double)sizeIndex};
    tf.write((const char *)hdr,sizeof(hdr));
    for (long long n=loLim[sizeIndex];n<=hiLim[sizeIndex];n++)
    {
      Eisenstein e=toFlowsnake((int)n);
      if (!tiles.count(e))
	continue;
      Tile &t=tiles[e];
      if (!t.nPoints)
	continue;
      int32_t iv[6]={(int32_t)n,e.getx(),e.gety(),t.nPoints,t.treeFlags,t.nGround};
      double dv[3]={t.density,t.hyperboloidSize,t.height};
      tf.write((const char *)iv,sizeof(iv));
      tf.write((const char *)dv,sizeof(dv));
      nTiles++;
    }
  }
  else
    fprintf(stderr,"could not identify flowsnake size index\n");
  xyz c=octRoot.getCenter();
  printf("{\"points\": %zu, \"stored\": %zu, \"bad_gpstime\": %zu, \"threads\": %d, \"canonical\": %s, "
	 "\"blocks\": %zu, \"duplicates\": %zu, \"root_center\": [%.17g, %.17g, %.17g], \"root_side\": %.17g, "
	 "\"snake_index\": %d, \"spacing\": %.17g, \"nonempty_tiles\": %zu, "
	 "\"read_build_s\": %.4f, \"canonical_s\": %.4f, \"scan_s\": %.4f, \"postscan_s\": %.4f, \"classify_s\": %.4f}\n",
	 total,nStored,nBadTime,nthreads,canonical?"true":"false",
	 (size_t)octStore.getNumBlocks(),alreadyInOctree.size(),c.getx(),c.gety(),c.getz(),octRoot.getSide(),
	 sizeIndex,snake.getSpacing(),nTiles,
	 tRead-t0,tCanon-tRead,tScan-tCanon,tPost-tScan,tClass-tPost);
  octStore.close();
  for (int i=0;i<nthreads+(int)relprime(nthreads);i++)
    remove((out+".store.oct"+to_string(i)).c_str());
  return 0;
}
