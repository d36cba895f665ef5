/* wb_oracle.h — CPU restatement of the reference's ground-extraction path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under wolkenbase_b200/ may include, link or call this;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs do.
 * Parity status: PINNED against the compiled, unmodified reference (oracle/_ref/ref_driver,
 * built from /root/reference by oracle/Makefile) and against the reference's own known-answer
 * tests (wolkentest.cpp:188-279, 744-773, 778-796) — see tests/test_oracle_*.py.
 */
#ifndef WB_ORACLE_H
#define WB_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define WBO_RECORDS 537          /* octree.h:41 */
#define WBO_LEVELS 21            /* depth of the canonical Morton key */

typedef struct wbo_tile
{
  int32_t n;                     /* flowsnake sequence number */
  int32_t ex,ey;                 /* Eisenstein address */
  int32_t nPoints,treeFlags;
  double density,hyperboloidSize,height;   /* tile.h:27-35 */
} wbo_tile;

typedef struct wbo_leaf
{
  uint64_t first;                /* first point (canonical order) */
  uint32_t count;
  int32_t depth;                 /* cube side = root side / 2^depth */
  double cx,cy,cz,half;          /* cube centre and half side */
} wbo_leaf;

/* --- angle.cpp / shape.cpp / flowsnake.cpp / leastsquares.cpp primitives (for KATs) --- */
void wbo_fill_tan_tables(void);
const double *wbo_tan_table(void);   /* 511 */
const double *wbo_cos_table(void);   /* 512 */
const double *wbo_sin_table(void);   /* 512 */
int wbo_atan2i(double y,double x);
int wbo_hyperboloid_in(const double v[3],double r,double s,const double p[3]);
int wbo_cylinder_in(double cx,double cy,double r,double px,double py);
int wbo_shape_in(int type,const double q[6],const double p[3]);
int wbo_shape_intersects_cube(int type,const double q[6],const double c[3],double side);
void wbo_shape_filter(int type,const double q[6],const double *pts,uint64_t n,uint8_t *in);
int wbo_cylinder_intersects_cube(double cx,double cy,double r,const double cube_center[3],double side);
void wbo_to_flowsnake(int n,int *ex,int *ey);
int wbo_from_flowsnake(int ex,int ey,int64_t *n);   /* our inverse; returns 0 if representable */
int wbo_base_seven(int ex,int ey);
double wbo_pairwise_sum(const double *a,unsigned n);
int wbo_least_squares(const double *a,const double *b,int rows,int cols,double *x);
int wbo_surround(const int32_t *dirs,int n);        /* 1 if the direction set surrounds */
int wbo_ldecimal(double x,char *buf,int buflen);

/* --- the path --- */
int wbo_decode(const uint8_t *recs,uint64_t n,int fmt,int rec_len,int32_t *xyz,uint8_t *cls,uint8_t *ret_num);
void wbo_coords(const int32_t *xyz,uint64_t n,const double scale[3],const double offset[3],double unit,double *out);
void wbo_size_fit(const double *corners,int n_corners,double center[3],double *side);
void wbo_bbox_cube(const double *corners,int n_corners,double cube[4]);
int wbo_snake_set_size(double cube_side,double tile_size,double *spacing,int *lo,int *hi);
uint64_t wbo_morton_key(const double p[3],const double center[3],double side);

/* canonical order = (21-level key, input index); order[k] = input index of k-th point */
int wbo_sort(const double *pts,uint64_t n,const double center[3],double side,uint64_t *keys_sorted,uint32_t *order);
/* leaves of the bucket octree (capacity 537), DFS order; returns the number of leaves (<= cap) */
int64_t wbo_leaves(const uint64_t *keys_sorted,uint64_t n,const double center[3],double side,wbo_leaf *out,int64_t cap);
int64_t wbo_dump(const wbo_leaf *leaves,int64_t n_leaves,char *buf,int64_t buflen);

/* scan + postscan over points given in canonical order.  Returns number of non-empty tiles. */
int64_t wbo_scan(const double *pts_sorted,uint64_t n,const double cube[4],double tile_size,
                 double min_hyperboloid_size,wbo_tile *out,int64_t cap);
int wbo_postscan(wbo_tile *tiles,int64_t n_tiles,double spacing);

/* per-point ground test; labels in the same (canonical) order as pts_sorted.  margin_count
 * receives the number of points with at least one in/out test within rel. 1e-12 of the surface. */
int wbo_classify(const double *pts_sorted,uint64_t n,const double cube[4],double tile_size,
                 double max_slope,double thickness,const wbo_tile *tiles,int64_t n_tiles,
                 uint8_t *labels,uint64_t *margin_count);

/* the same for the points at positions sel[0..n_sel) only, labels[k] for sel[k] (a sample of a very large cloud) */
int wbo_classify_sel(const double *pts_sorted,uint64_t n,const double cube[4],double tile_size,
                     double max_slope,double thickness,const wbo_tile *tiles,int64_t n_tiles,
                     const uint64_t *sel,uint64_t n_sel,uint8_t *labels,uint64_t *margin_count);

/* hyperboloidSize of the tile that classifies each point (NaN: in no tile); same arguments as wbo_classify */
int wbo_point_hyperboloid_sizes(const double *pts_sorted,uint64_t n,const double cube[4],double tile_size,
                                const wbo_tile *tiles,int64_t n_tiles,double *out);

#ifdef __cplusplus
}
#endif
#endif
