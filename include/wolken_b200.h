/* wolken_b200.h — C ABI of libwolken_b200.so, the B200 replacement for wolkenbase's worker pool.
 *
 * The reference runs its ground-extraction pipeline by flipping a thread-pool command
 * (threads.h:51-58: TH_READ, TH_SCAN, TH_POSTSCAN, TH_SPLIT, TH_PAUSE) and letting every
 * worker execute WolkenThread::operator() (threads.cpp:446-695).  This library replaces that
 * pool: each phase is one call that launches sm_100a kernels.  A reference maintainer binds
 * these entry points where the pool is driven today (INTEGRATION.md shows the shim):
 *
 *   phase (reference)                                   entry point here
 *   ------------------------------------------------    ---------------------------------
 *   startThreads(n)            threads.cpp:91-113        wb_create
 *   octRoot.sizeFit, snake.setSize, initTiles            wb_add_extent (+ implicit in wb_build)
 *        wolkencanvas.cpp:502-519, octree.cpp:268-310
 *   ACT_READ: readPoint + embufferPoint + OctStore::put  wb_add_las / wb_add_las_device + wb_build
 *        threads.cpp:477-590, las.cpp:735-820, octree.cpp:849-876
 *   octStore.dump              octree.cpp:888-891        wb_num_leaves / wb_get_leaves
 *   TH_SCAN: scanCylinder      scan.cpp:31-140           wb_scan
 *   TH_POSTSCAN: postscanCylinder  scan.cpp:142-179      wb_postscan
 *   TH_SPLIT: classifyCylinder classify.cpp:96-173       wb_classify
 *   ACT_COUNT: countClasses    threads.cpp:425-444       wb_count_classes
 *   ACT_WRITE: class byte of writePoint las.cpp:822-851  wb_get_labels / wb_patch_records
 *
 * Conventions: every function returns 0 on success and a negative code on failure (the
 * reference has no error codes: it asserts or prints); wb_last_error gives the
 * message.  Nothing throws.  All pointers are plain host pointers unless named d_*.  A context
 * is used from one host thread at a time.  There is no CPU fallback: without a CUDA device
 * wb_create fails.
 */
#ifndef WOLKEN_B200_H
#define WOLKEN_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct wb_ctx wb_ctx;

#define WB_RECORDS 537            /* leaf bucket capacity, octree.h:41 */
#define WB_LEVELS 21              /* depth of the canonical Morton key */

enum
{
  WB_OK=0,
  WB_ERR_CUDA=-1,
  WB_ERR_ARG=-2,
  WB_ERR_STATE=-3,
  WB_ERR_NOMEM=-4,
  WB_ERR_FORMAT=-5
};

typedef struct wb_leaf              /* one octree leaf = one OctBuffer, octree.h:95-144 */
{
  uint64_t first;                   /* index of its first point in canonical (Morton) order */
  uint32_t count;                   /* points in the bucket (<= 537 unless depth == WB_LEVELS) */
  int32_t depth;                    /* cube side = root side / 2^depth */
  double cx,cy,cz,half;             /* cube centre and half side (Octree::cube, octree.cpp:348-358) */
  double low,high;                  /* z range of the bucket (OctBuffer::low/high, octree.cpp:650-653) */
} wb_leaf;

typedef struct wb_tile              /* tile.h:27-35 */
{
  int32_t n;                        /* flowsnake sequence number (flowsnake.cpp:240-254) */
  int32_t ex,ey;                    /* Eisenstein address */
  int32_t nPoints,treeFlags;
  int32_t pad_;
  double density,hyperboloidSize,height;
} wb_tile;

typedef struct wb_geometry
{
  double root_center[3],root_side;  /* Octree::sizeFit result */
  double cube[4];                   /* bounding cube given to Flowsnake::setSize: cx,cy,cz,side */
  double spacing,radius;            /* tile spacing and cylinder radius (spacing*41/71) */
  int32_t snake_index,snake_lo,snake_hi,pad_;
} wb_geometry;

typedef struct wb_stats
{
  uint64_t n_points;                /* points in the store (after return-number-0 dropping) */
  uint64_t n_dropped;               /* records dropped by the return-number rule, threads.cpp:527-530 */
  uint64_t n_duplicates;            /* records whose XYZ equals an earlier point's: not stored (octree.cpp:620-662);
                                       they receive the stored point's class */
  uint64_t n_leaves;
  uint64_t n_tiles_nonempty;
  uint64_t n_memberships;           /* (point,tile) pairs = sum of tile nPoints */
  uint64_t n_margin;                /* points with an in/out test within rel. 1e-12 of the surface */
  uint64_t n_untiled;               /* points in no tile cylinder (left unclassified) */
  uint64_t n_second_walk;           /* points whose widest gap was 135..152 degrees: exact bearings needed */
  uint64_t cl_nodes,cl_chunks,cl_pairs; /* classify work: hierarchy nodes visited, chunks opened, (query,chunk) pairs tested */
  uint64_t cl_nodes2,cl_chunks2,cl_pairs2,cl_warps2; /* the same for the exact second walk, and warps that took it */
  uint64_t kernel_launches;         /* kernels launched by this context so far */
  double ms_h2d,ms_decode,ms_build,ms_scan,ms_postscan,ms_classify,ms_d2h;  /* last run, CUDA events */
  double ms_sort,ms_leaves,ms_hier,ms_pairs,ms_classify_kernel;
  double ms_encode,ms_encode_d2h;   /* wb_encode: kernel, copy of the records to the host */
  double ms_classify_order;         /* wb_classify: Hilbert re-sort of the store + hierarchy over it (part of ms_classify) */
} wb_stats;

/* ---- life cycle ---------------------------------------------------------- */
int wb_device_count(int *out);                      /* CUDA devices this process can use */
int wb_create(int device,wb_ctx **out);
void wb_destroy(wb_ctx *ctx);
const char *wb_last_error(wb_ctx *ctx);
int wb_reserve(wb_ctx *ctx,uint64_t n_points);     /* pre-size device arrays (optional) */
int wb_clear(wb_ctx *ctx);                          /* forget the cloud, keep the allocations */

/* tileSize, maxSlope, thickness, minimumHyperboloidSize: mainwindow.cpp:398-411 defaults 1,1,0,0.1 */
int wb_set_params(wb_ctx *ctx,double tile_size,double max_slope,double thickness,double min_hyperboloid_size);

/* ---- read + build (TH_READ / ACT_READ) ----------------------------------- */
/* Header corners of an input file (LasHeader::minCorner/maxCorner * unit). */
int wb_add_extent(wb_ctx *ctx,const double min_corner[3],const double max_corner[3]);
/* Point records of one file, host memory (pinned or pageable); copied in chunks and decoded
 * on the device as they arrive.  fmt 0-3 and 6-8 (las.cpp:38). */
int wb_add_las(wb_ctx *ctx,const uint8_t *recs,uint64_t n,int fmt,int rec_len,
               const double scale[3],const double offset[3],double unit);
/* ACT_READ drops records whose return number is 0 when the file's first record has a non-zero return number
 * (threads.cpp:485-500, 527-530) — the default here, decided per wb_add_las* call.  keep_all = 1 stores every
 * record, as a caller that feeds embufferPoint itself does (wolkencli.cpp:104-108). */
int wb_set_return_zero_rule(wb_ctx *ctx,int keep_all);
/* Of the records of every following wb_add_las* call keep those whose x = (offset+scale*X)*unit lies in [x_lo,x_hi),
 * in their order; the others are as if they were not in the file (return-number rule still decided by the file's
 * record 0).  For the ranks of a sharded run that are all handed the SAME whole files: rank r takes its x-interval
 * (wolkencli --gpus N with fewer files than GPUs).  -INFINITY, +INFINITY switches it off; wb_clear does too.  Not
 * together with wb_keep_records. */
int wb_set_window(wb_ctx *ctx,double x_lo,double x_hi);
int wb_num_loaded(wb_ctx *ctx,uint64_t *n);         /* records the context holds (all wb_add_las* calls so far) */
/* Same, straight from the file (LasHeader::readPoint's seek+read per point, las.cpp:735-745,
 * becomes a pipeline): worker threads pread 1 Mi-record chunks starting at byte point_offset into a
 * ring of pinned buffers while earlier chunks are copied and decoded.  A short file is an error. */
int wb_add_las_file(wb_ctx *ctx,const char *path,uint64_t point_offset,uint64_t n,int fmt,int rec_len,
                    const double scale[3],const double offset[3],double unit);
/* Same, records already in device memory. */
int wb_add_las_device(wb_ctx *ctx,const uint8_t *d_recs,uint64_t n,int fmt,int rec_len,
                      const double scale[3],const double offset[3],double unit);
/* Already-decoded points in device memory (SoA int32 X,Y,Z + class byte); they become one more segment with its
 * own scale/offset, every point kept (return number taken as 1). */
int wb_add_points_device(wb_ctx *ctx,const int32_t *d_x,const int32_t *d_y,const int32_t *d_z,const uint8_t *d_cls,
                         uint64_t n,const double scale[3],const double offset[3],double unit);
/* Only points whose input index lies in [first,end) are labelled by wb_classify; the others (context points, e.g. a
 * halo) take part in every query but keep label 255 — unless one of them holds the place of a record in the range
 * that has the same XYZ: that one is classified too, and the record inherits its class. */
int wb_set_own_range(wb_ctx *ctx,uint64_t first,uint64_t end);
/* Override the geometry derived from the extents (multi-GPU: every rank uses the global one). */
int wb_set_geometry(wb_ctx *ctx,const double root_center[3],double root_side,const double cube[4]);
int wb_get_geometry(wb_ctx *ctx,wb_geometry *out);
/* Morton keys, radix sort, leaf split, bucket hierarchy. */
int wb_build(wb_ctx *ctx);
int wb_num_leaves(wb_ctx *ctx,uint64_t *n);
int wb_get_leaves(wb_ctx *ctx,wb_leaf *out,uint64_t cap);
/* canonical order: order[k] = input index of the k-th point; keys[k] = its 63-bit Morton key */
int wb_get_order(wb_ctx *ctx,uint32_t *order,uint64_t *keys);
/* coordinates in canonical order, exactly the doubles the reference's LasPoint::location holds */
int wb_get_points_sorted(wb_ctx *ctx,double *x,double *y,double *z);
/* decoded SoA columns in input order (any pointer may be NULL) */
int wb_get_decoded(wb_ctx *ctx,int32_t *x,int32_t *y,int32_t *z,uint8_t *cls);

/* ---- scan / postscan (TH_SCAN, TH_POSTSCAN) ------------------------------ */
int wb_scan(wb_ctx *ctx);
int wb_postscan(wb_ctx *ctx);
int wb_num_tiles(wb_ctx *ctx,uint64_t *n);                 /* non-empty tiles */
int wb_get_tiles(wb_ctx *ctx,wb_tile *out,uint64_t cap);   /* ascending n */
/* Replace hyperboloidSize of the listed tiles (classify parity independent of scan parity). */
int wb_set_tiles(wb_ctx *ctx,const wb_tile *tiles,uint64_t n);

int wb_max_hyperboloid_size(wb_ctx *ctx,double *out);
/* Tile membership only (which tile's parameters each point uses); implied by wb_scan. */
int wb_assign(wb_ctx *ctx);

/* ---- classify (TH_SPLIT) -------------------------------------------------- */
int wb_classify(wb_ctx *ctx);
int wb_get_labels(wb_ctx *ctx,uint8_t *labels);            /* input order, one byte per record */
int wb_count_classes(wb_ctx *ctx,uint64_t counts[256]);
/* Write each point's class into its record (byte 15 low 5 bits for formats 0-5, byte 16 for
 * 6-10: las.cpp:754-756, 771, 848, 857) — host records, in place. */
int wb_patch_records(wb_ctx *ctx,uint8_t *recs,uint64_t first_point,uint64_t n,int fmt,int rec_len);

/* ---- several GPUs ---------------------------------------------------------------------
 * The analogue of startThreads(n) (threads.cpp:91-113) across GPUs: n ranks, one context each, every rank
 * holding the files of one x-strip of the cloud (ranks in ascending x).  wb_shard_run then does the whole
 * path for the rank's own records — global geometry from everybody's header corners, halo exchange for the
 * tile scan, the populated/tree grid for postscan, the wide halo for classify — and leaves the class bytes
 * of the rank's own records for wb_shard_get_labels.  Labels, tile parameters and canonical order equal the
 * single-GPU run's (csrc/wb_shard.cuh).
 *
 * A wb_comm is the rank's end of the exchange.  wb_comm_init: NCCL (one rank calls wb_comm_get_id and hands the
 * 128 bytes to the others, e.g. through the launcher); ranks may be processes or threads of one process.
 * wb_comm_init_local: the threads of ONE process without NCCL (plain copies + a barrier).  wb_comm_init_custom:
 * the caller moves the bytes (device pointers, whole-world collectives; return 0 on success). */
#define WB_COMM_ID_BYTES 128
#define WB_MAX_RANKS 64
#define WB_SHARD_MAXSEG 512         /* input files per rank */
#define WB_SHARD_SCAN_HALO 2.5      /* tile spacings sent across a strip border for the tile scan */
typedef struct wb_comm wb_comm;
typedef struct wb_local_group wb_local_group;
typedef struct wb_comm_ops
{
  void *user;
  /* every rank's `bytes` bytes at d_send, in rank order, into d_recv on every rank */
  int (*all_gather)(void *user,const void *d_send,void *d_recv,uint64_t bytes);
  /* send_cnt[k] bytes at d_send+send_off[k] go to rank k; recv_cnt[k] bytes from rank k arrive at d_recv+recv_off[k] */
  int (*all_to_all_v)(void *user,const void *d_send,const uint64_t *send_off,const uint64_t *send_cnt,
                      void *d_recv,const uint64_t *recv_off,const uint64_t *recv_cnt);
  /* element-wise maximum of n bytes over all ranks, in place */
  int (*all_reduce_max_u8)(void *user,uint8_t *d_buf,uint64_t n);
} wb_comm_ops;
typedef struct wb_shard_stats
{
  uint64_t n_own,own_first;                    /* own records; their first index in the rank's local input order */
  uint64_t n_halo_scan,n_halo_classify;        /* halo points received for the two stages */
  uint64_t grid_cells;                         /* cells of the all-reduced populated/tree grid */
  uint64_t bytes_sent,bytes_received;          /* halo rows, 16 B each */
  double por_max;                              /* largest hyperboloidSize * maxSlope^2 of the whole cloud */
  /* wall-clock milliseconds of the stages of the last wb_shard_run, stream drained at each boundary */
  double ms_setup,ms_select,ms_exchange,ms_build_scan,ms_scan,ms_grid,ms_postscan,ms_build_classify,ms_assign,ms_classify;
} wb_shard_stats;
int wb_comm_get_id(uint8_t id[WB_COMM_ID_BYTES]);
int wb_comm_init(wb_ctx *ctx,const uint8_t id[WB_COMM_ID_BYTES],int rank,int world,wb_comm **out);
int wb_local_group_create(int world,wb_local_group **out);
void wb_local_group_destroy(wb_local_group *grp);
int wb_comm_init_local(wb_ctx *ctx,wb_local_group *grp,int rank,wb_comm **out);
int wb_comm_init_custom(wb_ctx *ctx,const wb_comm_ops *ops,int rank,int world,wb_comm **out);
void wb_comm_destroy(wb_comm *comm);
/* after wb_add_extent + wb_add_las* of the rank's own files; collective: every rank calls it */
int wb_shard_run(wb_ctx *ctx,wb_comm *comm);
int wb_shard_get_labels(wb_ctx *ctx,uint8_t *labels);      /* the rank's own records, in the order they were added */
int wb_shard_get_stats(wb_ctx *ctx,wb_shard_stats *out);
/* Class bytes computed elsewhere (the ranks of a sharded run), input order, into a BUILT store: wb_count_classes,
 * wb_encode and the writers then work as after wb_classify (wolkencli --gpus N writes through one context). */
int wb_set_labels(wb_ctx *ctx,const uint8_t *labels);

/* ---- store queries -----------------------------------------------------------
 * OctStore::pointsIn / countPointsIn / hiLoPointsIn (octree.cpp:1214-1293) for the shapes of
 * shape.cpp, on the device: the exact Shape::in predicates (same operations as shape.cpp:81-90,
 * 127-135, 175-178, 215-219, 252-255) over the points a walk of the bounds hierarchy reaches.
 * Shape parameters are the constructors' arguments:
 *   WB_SPHERE      p = cx,cy,cz,radius            WB_PARABOLOID  p = vx,vy,vz,radiusCurvature
 *   WB_HYPERBOLOID p = vx,vy,vz,r,slope           WB_CYLINDER    p = cx,cy,radius
 *   WB_COLUMN      p = cx,cy,side
 * wb_query_batch: one result per shape (count; lowest and highest z, +inf/-inf when empty — the
 * per-pixel call of WolkenCanvas::pixelColorRead, wolkencanvas.cpp:92-108, for a whole raster).
 * wb_query_points: the points in one shape, in the order pointsIn returns them (bucket order x
 * in-bucket order); *n_out is the full count even when cap is smaller; any output may be NULL. */
enum { WB_SPHERE=0,WB_PARABOLOID=1,WB_HYPERBOLOID=2,WB_CYLINDER=3,WB_COLUMN=4 };
typedef struct wb_shape { int32_t type,pad_; double p[6]; } wb_shape;
int wb_query_batch(wb_ctx *ctx,const wb_shape *shapes,uint64_t n,uint64_t *count,double *lo,double *hi);
int wb_query_points(wb_ctx *ctx,const wb_shape *shape,uint64_t cap,uint64_t *n_out,uint32_t *pos,uint32_t *input_index,
                    double *x,double *y,double *z);

/* ---- output records (ACT_WRITE) ------------------------------------------------
 * CloudOutput::writeFiles (cloudoutput.cpp:187-229) walks the buckets in order and, for each class,
 * appends the bucket's points of that class to the class's least-full file through
 * LasHeader::writePoint (las.cpp:822-904): attributes as readPoint decoded them (las.cpp:735-820),
 * the new class, XYZ re-quantised as lrint((x/unit-offset)/scale).  Here the caller keeps the
 * sequential file choice (it needs only per-bucket class counts) and the device makes the records:
 *   wb_keep_records(ctx,1)            before wb_add_las: keep the raw records in device memory
 *   wb_leaf_class_counts              counts[leaf*K+k] = points of bucket `leaf` with class classes[k]
 *                                     (K = n_classes, or 1 and every class when separate == 0)
 *   wb_encode                         dest[leaf*K+k] = byte offset in `out` where that run of records
 *                                     starts (even), file_of[leaf*K+k] = index of the file it belongs
 *                                     to; out (may be NULL) receives all records, stats[f] the header figures of
 *                                     file f: per-return counts (n_points[0] = total) and the extremes
 *                                     of the written integers (header min/max = i*scale+offset).
 * Points whose class has no slot are not written (cloudoutput.cpp:212-214). */
typedef struct wb_out_spec
{
  int32_t format,rec_len;           /* output point format (0-3, 6-8) and its record length, las.cpp:38 */
  int32_t n_classes,separate;       /* CloudOutput::separateClasses */
  double scale[3],offset[3],unit;   /* LasHeader::setScale (las.cpp:636-673) result, file units */
  uint8_t classes[256];
} wb_out_spec;
typedef struct wb_file_stats
{
  uint64_t n_points[16];            /* LasHeader::nPoints: total, then by return number */
  int32_t imin[3],imax[3],pad_[2];
} wb_file_stats;
int wb_keep_records(wb_ctx *ctx,int keep);
int wb_leaf_class_counts(wb_ctx *ctx,const uint8_t *classes,int n_classes,int separate,uint32_t *counts);
int wb_encode(wb_ctx *ctx,const wb_out_spec *spec,const uint64_t *dest,const uint32_t *file_of,uint32_t n_files,
              uint8_t *out,uint64_t out_bytes,wb_file_stats *stats);
/* wb_encode with out == NULL leaves the records in device memory; wb_write_encoded streams the span
 * [arena_off, arena_off+bytes) of them into the open file descriptor at file_pos (D2H through a pinned
 * ring, pwrite on worker threads). */
int wb_write_encoded(wb_ctx *ctx,int fd,uint64_t file_pos,uint64_t arena_off,uint64_t bytes);
/* censusPoints() (testpattern.cpp:56-123, run after every write, threads.cpp:613): test data carries the point number
 * as GPS time; every point of the store sets one bit.  status -1: some GPS time is not a non-negative integer (not
 * test data; nothing else is reported), 1: a number occurs twice ("Duplicate point"), 0: neither.  max_point = one
 * past the highest number, n_missing = numbers below it that no stored point carries; up to `cap` of them,
 * ascending, go to `missing` (may be NULL).  Needs the records on the device (wb_keep_records) and a built store. */
typedef struct wb_census_result
{
  int32_t status,pad_;
  uint64_t n_stored,max_point,n_missing,n_duplicate;
} wb_census_result;
int wb_census(wb_ctx *ctx,wb_census_result *out,uint64_t *missing,uint64_t cap);
/* Records lost to an identical XYZ (octree.cpp:620-662): input index of each and of the point that
 * holds its place in the store; wb_stats.n_duplicates entries. */
int wb_get_duplicates(wb_ctx *ctx,uint32_t *dup,uint32_t *rep,uint64_t cap);

/* Device arithmetic exposed for known-answer tests (angle.cpp:117-155 atan2i, libm hypot as
 * point.cpp:189-192 uses it, the 64-sector binning of the classifier; -1 = "ask atan2i"). */
int wb_test_math(wb_ctx *ctx,uint64_t n,const double *y,const double *x,int32_t *atan2i_out,double *hypot_out,
                 int32_t *sector_out);

/* ---- whole pipeline -------------------------------------------------------- */
int wb_run(wb_ctx *ctx);                                    /* build, scan, postscan, classify */
int wb_get_stats(wb_ctx *ctx,wb_stats *out);
int wb_sync(wb_ctx *ctx);
/* Timing marks: wb_mark records CUDA event `slot` (0..WB_MARKS-1) on the stream the kernels are launched
 * on, behind all work issued so far; wb_mark_elapsed waits for mark `to` and returns the device time
 * between two marks.  (The reference times phases with wall clocks around waitForThreads,
 * wolkencanvas.cpp:341-354; a caller of this library cannot see its private stream otherwise.) */
#define WB_MARKS 8
int wb_mark(wb_ctx *ctx,int slot);
int wb_mark_elapsed(wb_ctx *ctx,int from,int to,double *ms);

/* ---- host helpers --------------------------------------------------------- */
int wb_host_alloc(void **p,uint64_t bytes);                 /* pinned host memory */
int wb_host_free(void *p);
/* host-side restatements used by the shims (no device work) */
int wb_size_fit(const double *corners,int n_corners,double center[3],double *side);
int wb_bbox_cube(const double *corners,int n_corners,double cube[4]);
int wb_bound_rect(const double *corners,int n_corners,double box[6]);   /* BoundRect: left,bottom,low,right,top,high */
int wb_snake_set_size(double cube_side,double tile_size,double *spacing,int *lo,int *hi);
int wb_ldecimal(double x,char *buf,int buflen);
int wb_format_dump(const wb_leaf *leaves,uint64_t n_leaves,char *buf,uint64_t buflen);

#ifdef __cplusplus
}
#endif
#endif
