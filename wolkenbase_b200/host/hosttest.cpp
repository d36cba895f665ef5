// hosttest — the parts of the reference-shaped C++ surface that need no GPU, checked the way the reference's
// wolkentest checks its own (assertions, exit status): cloud.cpp's block view of the non-LAS point list
// (cloud.cpp:27-45) and testpattern.cpp's per-block census (testpattern.cpp:56-82).
#include <cassert>
#include <cmath>
#include <iostream>
#include "wolken_host.h"

using namespace std;

static LasPoint numbered(double t)
{
  LasPoint p;
  p.location=xyz(1,2,3);
  p.gpsTime=t;
  return p;
}

int main()
{
  // ---- cloud: RECORDS points per block, the last one ragged, default attributes
  assert(getNumCloudBlocks()==0);
  assert(getCloudBlock(0).empty());
  const size_t n=2*RECORDS+17;
  for (size_t i=0;i<n;i++)
    cloud.push_back(xyz((double)i,2.0*i,-1.0*i));
  assert(getNumCloudBlocks()==3);
  assert(getCloudBlock(0).size()==RECORDS && getCloudBlock(1).size()==RECORDS && getCloudBlock(2).size()==17);
  assert(getCloudBlock(3).empty() && getCloudBlock(-1).empty());
  vector<LasPoint> b=getCloudBlock(2);
  assert(b[0].location.getx()==(double)(2*RECORDS) && b[16].location.getz()==-1.0*(n-1));
  assert(b[5].classification==0);
  cloud.clear();
  assert(getNumCloudBlocks()==0);
  // ---- census of one block: 0 = all new, 1 = a number seen before, -1 = not test data
  vector<LasPoint> blk;
  for (int i=0;i<100;i++)
    blk.push_back(numbered(i));
  assert(censusPoints(blk)==0);
  assert(censusPoints(vector<LasPoint>(1,numbered(42)))==1);
  assert(censusPoints(vector<LasPoint>(1,numbered(100)))==0);
  assert(censusPoints(vector<LasPoint>(1,numbered(7.5)))==-1);
  assert(censusPoints(vector<LasPoint>(1,numbered(-3)))==-1);
  assert(censusPoints(vector<LasPoint>(1,numbered(NAN)))==-1);
  cout<<"hosttest ok\n";
  return 0;
}
