/*
 * rt_oracle.h — TEST INFRASTRUCTURE ONLY (see rt_oracle.c).  CPU restatement of the
 * reference's image-formation path, used as the parity checker for the CUDA path.
 * It shares the POD problem description of include/rtb200.h so that the checker and the
 * product are driven with the very same host arrays.
 */
#ifndef RT_ORACLE_H
#define RT_ORACLE_H

#include "../include/rtb200.h"

#ifdef __cplusplus
extern "C" {
#endif

size_t rt_oracle_findfirstsingle(const double *X, size_t size_X, double Y);
uint32_t rt_oracle_findindex(const double *X, uint32_t size_X, double Y);
float rt_oracle_bilinear(float dx, float dy, float f1, float f2, float f3, float f4);
double rt_oracle_interp_pchip(size_t N, const double *xi, const double *yi, double x);
void rt_oracle_calc_seed(const rtb200_seed *seed, double x, double y, double a, double b,
                         double *Iv);

int rt_oracle_calc_ray(const rtb200_ray *ray, int N, float dz0, const rtb200_gain_plane *gain,
                       const rtb200_seed *seed, int K, int method, float c, double *Iv,
                       rtb200_ray *ray2, float *gvl, float *evl, int32_t *ivl,
                       int *escaped_out, uint64_t *steps);

int rt_oracle_calc_ray_debug(const rtb200_ray *ray, int N, float dz0,
                             const rtb200_gain_plane *gain, const rtb200_seed *seed, int K,
                             int method, float c, const double *dv, double *Iv, rtb200_ray *ray2,
                             float *debug);

void rt_oracle_trace_rays(int N, const rtb200_beam *beam, const rtb200_gain_plane *gain,
                          const rtb200_seed *seed, int method, const rtb200_ray *rays,
                          size_t n_rays, double scale, double *image, double *I_ang,
                          unsigned *failure_code, rtb200_ray *failed, int max_failed,
                          int *n_failed, uint64_t *steps);

size_t rt_oracle_ray_count(const rtb200_problem *p);
void rt_oracle_build_rays(const rtb200_problem *p, rtb200_ray *rays, int *method, double *scale);
int rt_oracle_create_image(const rtb200_problem *p, unsigned flags, double *image, double *I_ang,
                           unsigned *failure_code, rtb200_ray *failed, int max_failed,
                           int *n_failed, uint64_t *steps);

/* Multi-threaded driver over rt_oracle_trace_rays (contiguous ray chunks, private partial
 * images summed on the host): the restatement of RayTraceImageThreadLoop
 * (src/RayTraceImage.cpp:89-134), used only as the timed host-CPU baseline. */
int rt_oracle_create_image_threads(const rtb200_problem *p, unsigned flags, int n_threads,
                                   double *image, double *I_ang,
                                   unsigned *failure_code, size_t *rays_done);

#ifdef __cplusplus
}
#endif
#endif
