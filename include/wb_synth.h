/* wb_synth.h — C ABI of the deterministic synthetic-cloud generator (bench / test input only).
 * Models the reference's testpattern.cpp scenes (laserize, testpattern.cpp:166-190) for the
 * five BASELINE.json configs; see wolkenbase_b200/csrc/synth.c. */
#ifndef WB_SYNTH_H
#define WB_SYNTH_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct wb_synth_desc
{
  int32_t scene;          /* 1 street, 2 aerial, 3 aerial in format 6, 4 terrestrial, 5 urban */
  int32_t fmt;            /* LAS point format written */
  double scale;           /* metres per integer tick (all three axes) */
  double offset[3];       /* LAS header offsets */
  uint64_t n_points;      /* points actually generated (grid_nx*grid_ny, x24 for scene 4) */
  uint64_t grid_nx,grid_ny;
  uint64_t cell_ticks;    /* jitter cell edge in ticks (0 for scene 4) */
  uint64_t extent_ticks;  /* scene edge in ticks */
} wb_synth_desc;

int wb_synth_record_length(int fmt);
int wb_synth_describe(int scene,uint64_t n_points,wb_synth_desc *d);
int wb_synth_generate(const wb_synth_desc *d,uint64_t seed,
                      uint64_t cell_x0,uint64_t cell_y0,uint64_t ncx,uint64_t ncy,
                      uint64_t gps_base,uint8_t *recs,int32_t bbox[6]);
int wb_synth_generate_terrestrial(const wb_synth_desc *d,uint64_t seed,uint64_t gps_base,
                                  uint8_t *recs,int32_t bbox[6]);
int wb_synth_header(const wb_synth_desc *d,uint64_t n_points,const int32_t bbox[6],uint8_t *hdr);

#ifdef __cplusplus
}
#endif
#endif
