// wolkenquery — exercises the OctStore query surface (findBlock, findBlocks, pointsIn,
// countPointsIn, hiLoPointsIn; octree.cpp:199-251, 1214-1293) on a LAS file, like the reference's
// testflat (wolkentest.cpp:143-167) does for a cylinder.  Used by tests/test_host_cli.py.
//   wolkenquery in.las cyl cx cy r | sph cx cy cz r | hyp vx vy vz r s | par vx vy vz rc | col cx cy side [--host]
// Queries run on the GPU (wb_query_*); --host runs the CPU mirror of the octree walk instead (the checker).
#include <cstdlib>
#include <cstring>
#include <deque>
#include <iostream>
#include "wolken_host.h"
using namespace std;

int main(int argc,char **argv)
{
  if (argc<4)
    return 2;
  if (!strcmp(argv[argc-1],"--host"))
  {
    hostQueries=true;
    argc--;
  }
  deque<LasHeader> files(1);
  files[0].openRead(argv[1]);
  if (!files[0].isValid())
    return 1;
  vector<xyz> limits={files[0].minCorner(),files[0].maxCorner()};
  octRoot.sizeFit(limits);
  vector<double> c;
  for (auto &p:limits) { c.push_back(p.getx()); c.push_back(p.gety()); c.push_back(p.getz()); }
  double cube[4];
  wb_bbox_cube(c.data(),2,cube);
  snake.setSize(Cube(xyz(cube[0],cube[1],cube[2]),cube[3]),tileSize);
  startThreads(1);
  waitForThreads(TH_READ);
  ThreadAction ta;
  ta.opcode=ACT_READ;
  ta.hdr=&files[0];
  enqueueAction(ta);
  waitForQueueEmpty();
  Shape *sh=nullptr;
  string kind=argv[2];
  if (kind=="cyl" && argc>=6)
    sh=new Cylinder(xy(atof(argv[3]),atof(argv[4])),atof(argv[5]));
  else if (kind=="sph" && argc>=7)
    sh=new Sphere(xyz(atof(argv[3]),atof(argv[4]),atof(argv[5])),atof(argv[6]));
  else if (kind=="hyp" && argc>=8)
    sh=new Hyperboloid(xyz(atof(argv[3]),atof(argv[4]),atof(argv[5])),atof(argv[6]),atof(argv[7]));
  else if (kind=="par" && argc>=7)
    sh=new Paraboloid(xyz(atof(argv[3]),atof(argv[4]),atof(argv[5])),atof(argv[6]));
  else if (kind=="col" && argc>=6)
    sh=new Column(xy(atof(argv[3]),atof(argv[4])),atof(argv[5]));
  else
    return 2;
  vector<int64_t> blocks=octRoot.findBlocks(*sh);
  vector<LasPoint> pts=octStore.pointsIn(*sh,true);
  array<double,2> hl=octStore.hiLoPointsIn(*sh);
  bool sorted=true,consistent=true;
  for (size_t i=1;i<pts.size();i++)
    sorted=sorted && pts[i-1].location.getz()<=pts[i].location.getz();
  for (auto &p:pts)
  {
    int64_t b=octRoot.findBlock(p.location);
    consistent=consistent && b>=0 && octRoot.findCube(p.location).in(p.location);
  }
  double gsum=0;
  for (auto &p:pts)
    gsum+=p.gpsTime;
  cout.precision(17);
  cout<<"{\"blocks\": "<<blocks.size()<<", \"count\": "<<octStore.countPointsIn(*sh)<<", \"points\": "<<pts.size()
      <<", \"lo\": "<<(pts.empty()?0:hl[0])<<", \"hi\": "<<(pts.empty()?0:hl[1])<<", \"sorted\": "<<sorted
      <<", \"consistent\": "<<consistent<<", \"gps_sum\": "<<gsum<<", \"total_blocks\": "<<octStore.getNumBlocks()
      <<", \"total_points\": "<<octStore.countPoints()<<"}"<<endl;
  return 0;
}
