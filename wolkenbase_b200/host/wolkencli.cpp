// wolkencli — the reference's command line (wolkencli.cpp:43-129: positional LAS inputs, header
// lines, octree cube, "dumpfile") driving the GPU library through the reference-shaped C++
// surface, extended with the GUI's classify-and-write sequence (wolkencanvas.cpp:469-625):
//   wolkencli [options] input.las...
//     -o, --output NAME          classify and write NAME[-class][-k].las
//     --tile-size T  --max-slope S  --thickness K  --min-hyperboloid-size M   (QSettings keys, mainwindow.cpp:398-411)
//     --points-per-file N        split outputs every N points (0 = no split)
//     --separate-classes 0|1     one file per class (default 1, as the GUI)
//     --gpus N, --threads N      spread scan, postscan and classify over N GPUs (startThreads(N), threads.cpp:91-113):
//                                with at least N files they are taken in ascending x and dealt out as N x-strips; with
//                                fewer (one big file) every GPU reads every file and keeps its x-interval
//     --census                   after writing, count the stored points by their GPS time (test data carries the point
//                                number there; censusPoints, threads.cpp:613): always done when the records are on
//                                the GPU (the default writer), on request with --lossless / --host-writer
//     --dump FILE                where to write the octree dump (default "dumpfile")
//     --host-writer              make the output records on the CPU (LasHeader::writePoint) instead of on the GPU
//     --timing                   print the wall time of each phase (seconds) as one JSON line at the end
//     --embuffer                 feed the store the way the reference's wolkencli does (readPoint + embufferPoint per
//                                point, wolkencli.cpp:104-108) instead of one ACT_READ per file
//     --lossless                 keep the inputs' own records (format, scale, offset) and only replace the class
//                                byte, instead of the reference's LAS 1.4 re-encoding (CloudOutput)
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <iostream>
#include <fstream>
#include <chrono>
#include "wolken_host.h"

using namespace std;

int main(int argc,char **argv)
{
  vector<string> inputFiles;
  OutputOptions out;
  string dumpName="dumpfile";
  bool classify=false,lossless=false,hostWriter=false,timing=false,perPoint=false,census=false;
  int nGpus=1;
  auto now=[]{ return chrono::duration<double>(chrono::steady_clock::now().time_since_epoch()).count(); };
  double t0=now(),tRead=0,tScan=0,tPost=0,tClass=0,tWrite=0,tDump=0,tOpen=0;
  for (int i=1;i<argc;i++)
  {
    string a=argv[i];
    auto val=[&]()->const char * { if (i+1>=argc) { cerr<<a<<" needs a value\n"; exit(2); } return argv[++i]; };
    if (a=="-o" || a=="--output") { out.baseName=val(); classify=true; }
    else if (a=="--tile-size") tileSize=atof(val());
    else if (a=="--max-slope") maxSlope=atof(val());
    else if (a=="--thickness") thickness=atof(val());
    else if (a=="--min-hyperboloid-size") minHyperboloidSize=atof(val());
    else if (a=="--points-per-file") out.pointsPerFile=strtoull(val(),nullptr,10);
    else if (a=="--separate-classes") out.separateClasses=atoi(val())!=0;
    else if (a=="--threads" || a=="--gpus") nGpus=max(1,atoi(val()));
    else if (a=="--dump") dumpName=val();
    else if (a=="--lossless") lossless=true;
    else if (a=="--host-writer") hostWriter=true;
    else if (a=="--timing") timing=true;
    else if (a=="--embuffer") perPoint=true;
    else if (a=="--census") census=true;
    else if (a.size() && a[0]=='-') { cerr<<"unknown option "<<a<<endl; return 2; }
    else inputFiles.push_back(a);
  }
  if (out.baseName.size()>4 && out.baseName.substr(out.baseName.size()-4)==".las")
    out.baseName.resize(out.baseName.size()-4);
  keepRecordsOnDevice=classify && !lossless && !hostWriter;
  if (nGpus>1 && inputFiles.size()>1)
  {
    // strips in ascending x: the order of the files is the input order of the points, on one GPU as on several
    vector<pair<double,string>> byX;
    for (auto &name:inputFiles)
    {
      LasHeader h;
      h.openRead(name);
      byX.emplace_back(h.isValid()?h.minCorner().getx():INFINITY,name);
    }
    stable_sort(byX.begin(),byX.end(),[](const pair<double,string> &a,const pair<double,string> &b){ return a.first<b.first; });
    for (size_t i=0;i<byX.size();i++)
      inputFiles[i]=byX[i].second;
  }
  deque<LasHeader> files(inputFiles.size());
  vector<xyz> limits;
  double mn[3]={INFINITY,INFINITY,INFINITY},mx[3]={-INFINITY,-INFINITY,-INFINITY};
  for (size_t i=0;i<inputFiles.size();i++)
  {
    files[i].openRead(inputFiles[i]);
    if (!files[i].isValid())
    {
      cerr<<inputFiles[i]<<": not a LAS file this program can read\n";
      return 1;
    }
    limits.push_back(files[i].minCorner());
    limits.push_back(files[i].maxCorner());
    int ver=files[i].getVersion();
    cout<<"Version "<<(ver>>8)<<'.'<<(ver&255)<<' ';
    cout<<files[i].numberPoints()<<" points, format "<<files[i].getPointFormat()<<endl;
  }
  if (files.empty())
  {
    cerr<<"usage: wolkencli [options] input.las...\n";
    return 2;
  }
  octRoot.sizeFit(limits);
  xyz center=octRoot.getCenter();
  char b0[64],b1[64],b2[64];
  wb_ldecimal(center.getx(),b0,64); wb_ldecimal(center.gety(),b1,64); wb_ldecimal(center.getz(),b2,64);
  cout<<'('<<b0<<','<<b1<<','<<b2<<")±"<<octRoot.getSide()<<endl;
  // the cube handed to the flowsnake: bounding box of the header corners (wolkencanvas.cpp:502-519)
  {
    vector<double> c;
    for (auto &p:limits) { c.push_back(p.getx()); c.push_back(p.gety()); c.push_back(p.getz()); }
    double cube[4];
    wb_bbox_cube(c.data(),(int)limits.size(),cube);
    snake.setSize(Cube(xyz(cube[0],cube[1],cube[2]),cube[3]),tileSize);
    initTiles();
    (void)mn; (void)mx;
  }
  tOpen=now()-t0;
  double t=now();
  startThreads(nGpus);
  waitForThreads(TH_READ);
  for (size_t i=0;i<files.size();i++)
  {
    if (perPoint)
      for (size_t j=0;j<files[i].numberPoints();j++)
        embufferPoint(files[i].readPoint(j),true);
    else
    {
      ThreadAction ta;
      ta.opcode=ACT_READ;
      ta.hdr=&files[i];
      enqueueAction(ta);
    }
    cout<<files[i].numberPoints()<<" points, "<<pointBufferSize()<<" points in buffer\n";
  }
  waitForQueueEmpty();
  cout<<"All points in octree\n";
  cout<<duplicatePoints()<<" duplicate points\n";
  cout<<octStore.getNumBlocks()<<" blocks\n";
  tRead=now()-t;
  if (classify)
  {
    // each transition runs the phase it names on the GPU and returns when it is done
    t=now();
    waitForThreads(TH_SCAN);
    tScan=now()-t; t=now();
    cout<<"Starting scan\n";
    waitForThreads(TH_POSTSCAN);
    tPost=now()-t; t=now();
    cout<<"Starting postscan\n";
    waitForThreads(TH_SPLIT);
    tClass=now()-t; t=now();
    cout<<"Starting classifying\n";
    waitForThreads(TH_PAUSE);
    cout<<"Counting points\n";
    classTotals.clear();
    ThreadAction ta;
    ta.opcode=ACT_COUNT;
    enqueueAction(ta);
    cout<<"Classified points:\n";
    for (auto &j:classTotals)
      cout<<j.first<<' '<<j.second<<endl;
    vector<string> written;
    int rc=lossless?writeClassified(files,out,&written):writeReferenceStyle(files,out,&written);
    if (rc)
      return 5;
    for (auto &w:written)
      cout<<"Wrote "<<w<<endl;
    if (census || keepRecordsOnDevice)
      censusPoints();
    tWrite=now()-t;
  }
  t=now();
  waitForThreads(TH_STOP);
  cout<<"Dumping octree\n";
  {
    ofstream dumpFile(dumpName);
    octStore.dump(dumpFile);
  }
  joinThreads();
  tDump=now()-t;
  wb_stats gst;
  memset(&gst,0,sizeof(gst));
  wb_get_stats(wolkenContext(),&gst);
  if (shardReport.world>1)
  {
    cout<<shardReport.world<<" GPUs, "<<shardReport.seconds<<" s:\n";
    for (int r=0;r<shardReport.world;r++)
    {
      const wb_shard_stats &q=shardReport.ranks[r];
      cout<<"  GPU "<<shardReport.device[r]<<": "<<shardReport.files[r]<<" files, "<<q.n_own<<" points, halo "<<q.n_halo_scan
          <<" + "<<q.n_halo_classify<<", classify "<<q.ms_classify<<" ms\n";
    }
  }
  if (timing)
    cout<<"{\"open_s\": "<<tOpen<<", \"read_build_s\": "<<tRead<<", \"scan_s\": "<<tScan<<", \"postscan_s\": "<<tPost
        <<", \"classify_s\": "<<tClass<<", \"count_write_s\": "<<tWrite<<", \"dump_s\": "<<tDump
        <<", \"wb_create_s\": "<<hostTimes[0]<<", \"wb_add_las_file_s\": "<<hostTimes[1]<<", \"wb_build_s\": "<<hostTimes[2]
        <<", \"wb_encode_write_s\": "<<hostTimes[3]<<", \"encode_kernel_ms\": "<<gst.ms_encode<<", \"total_s\": "<<now()-t0<<"}\n";
  return 0;
}
