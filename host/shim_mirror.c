/* shim_mirror.c -- the calls haskell/PzGpu.hs makes, in C99, in the same order and with the same buffer
 * arithmetic: descriptors allocated as sizeofModelDesc opaque bytes and patched at the byte offsets the Haskell
 * module pokes, the summary read at its byte offsets, particle rows chunked by the nx read from the descriptor,
 * track maps unpacked from the four read_tracks arrays.  GHC is not in this image, so this file is what runs in
 * CI in place of the Haskell module (tests/test_host_cli.py); it prints one JSON object per scenario.
 *
 *   shim_mirror [n_devices]      n_devices > 1 adds the withDevices scenario (pz_filter_create_multi)
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pzgpu.h"

#define SIZEOF_MODEL_DESC 1152 /* PzGpu.hs: sizeofModelDesc */
#define SIZEOF_SUMMARY 144     /* PzGpu.hs: sizeofSummary   */

static void check(int rc, const char* what) {
  if (rc != 0) { fprintf(stderr, "shim_mirror: %s: %d: %s\n", what, rc, pz_last_error()); exit(1); }
}

static int32_t desc_nx(const void* d) { int32_t v; memcpy(&v, (const char*)d + 4, 4); return v; }          /* descNx: peekByteOff p 4 */
static void poke_i32(void* d, int off, int32_t v) { memcpy((char*)d + off, &v, 4); }

/* readSummary: step @0, log_evidence_inc @8, ess @16, mean @24, var @72 */
static void print_summary(const char* ps, int nx) {
  int64_t st; double le, es, mu[6], va[6];
  memcpy(&st, ps + 0, 8); memcpy(&le, ps + 8, 8); memcpy(&es, ps + 16, 8);
  memcpy(mu, ps + 24, 8 * (size_t)nx); memcpy(va, ps + 72, 8 * (size_t)nx);
  printf("{\"step\":%lld,\"le\":%.17g,\"ess\":%.17g,\"mean\":[", (long long)st, le, es);
  for (int d = 0; d < nx; ++d) printf("%s%.17g", d ? "," : "", mu[d]);
  printf("],\"var\":[");
  for (int d = 0; d < nx; ++d) printf("%s%.17g", d ? "," : "", va[d]);
  printf("]}");
}

/* inferWith + stepf + particles k (PzGpu.hs) */
static void scenario_lg(const char* name, void* desc, int n, int n_devices, int stride) {
  pz_filter* f = NULL;
  if (n_devices > 1) {
    int ids[32];
    for (int i = 0; i < n_devices; ++i) ids[i] = i;
    check(pz_filter_create_multi((const pz_model_desc*)desc, (uint64_t)n, 7, n_devices, ids, &f), "pz_filter_create_multi");
  } else {
    check(pz_filter_create((const pz_model_desc*)desc, (uint64_t)n, 7, 0, &f), "pz_filter_create");
  }
  const int nx = desc_nx(desc);
  int32_t ny; memcpy(&ny, (const char*)desc + 8, 4);
  printf("{\"scenario\":\"%s\",\"nx\":%d,\"steps\":[", name, nx);
  char ps[SIZEOF_SUMMARY];
  for (int t = 0; t < 4; ++t) {
    double obs[3] = {0.3 * (t + 1), -0.2 * t, 0.1};
    check(pz_filter_step(f, obs, 1, ny, (pz_step_summary*)ps), "pz_filter_step");
    if (t) printf(",");
    print_summary(ps, nx);
  }
  printf("]");
  if (n_devices <= 1) {   /* particles k: rows = (n_local + k - 1) / k, each nx wide */
    const int64_t nl = pz_filter_n_local(f);
    const int64_t rows = (nl + stride - 1) / stride;
    double* buf = (double*)malloc(sizeof(double) * (size_t)(rows * nx));
    check(pz_filter_read_particles(f, stride, buf, rows * nx), "pz_filter_read_particles");
    printf(",\"rows\":%lld,\"first_particle\":[", (long long)rows);
    for (int d = 0; d < nx; ++d) printf("%s%.17g", d ? "," : "", buf[d]);
    printf("],\"last_particle\":[");
    for (int d = 0; d < nx; ++d) printf("%s%.17g", d ? "," : "", buf[(rows - 1) * nx + d]);
    printf("]");
    free(buf);
  } else {
    printf(",\"n_local\":%lld", (long long)pz_filter_n_local(f));
  }
  printf("}\n");
  pz_filter_destroy(f);
}

int main(int argc, char** argv) {
  const int n_devices = argc > 1 ? atoi(argv[1]) : 1;
  void* d = malloc(SIZEOF_MODEL_DESC);

  check(pz_model_rw1d((pz_model_desc*)d, 1.0, 3.0), "pz_model_rw1d");             /* rw1d 1 3 */
  scenario_lg("rw1d", d, 10000, 1, 1000);
  poke_i32(d, 24, 1);                                                              /* withDouble: state_f64 @24 */
  scenario_lg("rw1d_double", d, 10000, 1, 1000);

  check(pz_model_mtt_track((pz_model_desc*)d), "pz_model_mtt_track");              /* mttTrack */
  scenario_lg("mtt_track", d, 10000, 1, 999);

  check(pz_model_stt((pz_model_desc*)d), "pz_model_stt");                          /* stt */
  scenario_lg("stt", d, 10000, 1, 1000);
  poke_i32(d, 20, 1);                                                              /* sttDelayed: delayed @20 */
  scenario_lg("stt_delayed", d, 10000, 1, 5000);

  check(pz_model_walker((pz_model_desc*)d), "pz_model_walker");                    /* walker */
  scenario_lg("walker", d, 10000, 1, 1000);

  /* runAllGPU: pz_filter_run over the whole list, summaries in one array of sizeofSummary records */
  {
    check(pz_model_rw1d((pz_model_desc*)d, 1.0, 3.0), "pz_model_rw1d");
    pz_filter* f = NULL;
    check(pz_filter_create((const pz_model_desc*)d, 10000, 7, 0, &f), "pz_filter_create");
    double obs[4] = {0.3, 0.6, 0.9, 1.2};
    char* ps = (char*)malloc(4 * SIZEOF_SUMMARY);
    check(pz_filter_run(f, obs, 4, 1, (pz_step_summary*)ps), "pz_filter_run");
    printf("{\"scenario\":\"run_all\",\"nx\":1,\"steps\":[");
    for (int t = 0; t < 4; ++t) { if (t) printf(","); print_summary(ps + t * SIZEOF_SUMMARY, 1); }
    printf("]}\n");
    free(ps);
    pz_filter_destroy(f);
  }

  /* betaBernoulli 1 1 + read_posterior */
  {
    check(pz_model_beta_bernoulli((pz_model_desc*)d, 1.0, 1.0), "pz_model_beta_bernoulli");
    pz_filter* f = NULL;
    check(pz_filter_create((const pz_model_desc*)d, 100, 0, 0, &f), "pz_filter_create");
    char ps[SIZEOF_SUMMARY];
    double ys[3] = {1.0, 1.0, 0.0};
    for (int t = 0; t < 3; ++t) check(pz_filter_step(f, &ys[t], 1, 1, (pz_step_summary*)ps), "pz_filter_step");
    double mean[6], cov[36];
    check(pz_filter_read_posterior(f, mean, cov), "pz_filter_read_posterior");
    printf("{\"scenario\":\"beta_bernoulli\",\"alpha\":%.17g,\"beta\":%.17g}\n", mean[0], mean[1]);
    pz_filter_destroy(f);
  }

  /* mttGPU: variable-length detection lists, then tracks k */
  {
    check(pz_model_mtt((pz_model_desc*)d), "pz_model_mtt");
    pz_filter* f = NULL;
    check(pz_filter_create((const pz_model_desc*)d, 4000, 3, 0, &f), "pz_filter_create");
    char ps[SIZEOF_SUMMARY];
    double dets[3][4] = {{0.5, -0.4, 2.0, 1.0}, {0.6, -0.3, 0.0, 0.0}, {0.7, -0.2, -1.5, 2.5}};
    int cnt[3] = {2, 1, 2};
    for (int t = 0; t < 3; ++t) check(pz_filter_step(f, dets[t], cnt[t], 2, (pz_step_summary*)ps), "pz_filter_step (mtt)");
    const int k = 500;
    const int64_t nl = pz_filter_n_local(f), rows = (nl + k - 1) / k;
    int32_t* pc = (int32_t*)malloc(sizeof(int32_t) * (size_t)rows);
    int32_t* pid = (int32_t*)malloc(sizeof(int32_t) * (size_t)rows * 16);
    double* pm = (double*)malloc(sizeof(double) * (size_t)rows * 16 * 4);
    double* pv = (double*)malloc(sizeof(double) * (size_t)rows * 16 * 16);
    check(pz_filter_read_tracks(f, k, pc, pid, pm, pv, rows), "pz_filter_read_tracks");
    printf("{\"scenario\":\"mtt\",\"rows\":%lld,\"particles\":[", (long long)rows);
    for (int64_t i = 0; i < rows; ++i) {
      printf("%s[", i ? "," : "");
      for (int j = 0; j < pc[i]; ++j)
        printf("%s[%d,[%.9g,%.9g],%.9g]", j ? "," : "", pid[i * 16 + j], pm[(i * 16 + j) * 4], pm[(i * 16 + j) * 4 + 1], pv[(i * 16 + j) * 16]);
      printf("]");
    }
    printf("]}\n");
    free(pc); free(pid); free(pm); free(pv);
    pz_filter_destroy(f);
  }

  /* multiCameraGPU: observation rows (camera, pose[10], box[4], reading), then marginal tracks k (no appearance readout) */
  {
    check(pz_model_multicam((pz_model_desc*)d), "pz_model_multicam");
    pz_filter* f = NULL;
    check(pz_filter_create((const pz_model_desc*)d, 4000, 3, 0, &f), "pz_filter_create");
    char ps[SIZEOF_SUMMARY];
    double rows[2][16];
    for (int t = 0; t < 3; ++t) {
      for (int c = 0; c < 2; ++c) {
        for (int i = 0; i < 16; ++i) rows[c][i] = 0.0;
        rows[c][0] = c; rows[c][1 + (t + c) % 10] = 1.0;                       /* a unit pose */
        rows[c][11] = 0.3 * t - 0.2 * c; rows[c][12] = 0.1 * c; rows[c][13] = 1.0; rows[c][14] = 1.1; rows[c][15] = 0.4 - 0.3 * c;
      }
      check(pz_filter_step(f, &rows[0][0], 2, 16, (pz_step_summary*)ps), "pz_filter_step (multicam)");
    }
    const int k = 500;
    const int64_t nl = pz_filter_n_local(f), sel = (nl + k - 1) / k;
    int32_t* pc = (int32_t*)malloc(sizeof(int32_t) * (size_t)sel);
    int32_t* pid = (int32_t*)malloc(sizeof(int32_t) * (size_t)sel * 16);
    int32_t* pidn = (int32_t*)malloc(sizeof(int32_t) * (size_t)sel * 16);
    int32_t* pcam = (int32_t*)malloc(sizeof(int32_t) * (size_t)sel * 16);
    double* pst = (double*)malloc(sizeof(double) * (size_t)sel * 16);
    double* pm = (double*)malloc(sizeof(double) * (size_t)sel * 16 * 4);
    double* pv = (double*)malloc(sizeof(double) * (size_t)sel * 16 * 4);
    check(pz_filter_read_mcam_tracks(f, k, pc, pid, pidn, pcam, pst, pm, pv, NULL, NULL, NULL, sel), "pz_filter_read_mcam_tracks");
    printf("{\"scenario\":\"multicam\",\"rows\":%lld,\"particles\":[", (long long)sel);
    for (int64_t i = 0; i < sel; ++i) {
      printf("%s[", i ? "," : "");
      for (int j = 0; j < pc[i]; ++j)
        printf("%s[%d,%d,%d,[%.9g,%.9g],%.9g]", j ? "," : "", pid[i * 16 + j], pidn[i * 16 + j], pcam[i * 16 + j], pm[(i * 16 + j) * 4],
               pm[(i * 16 + j) * 4 + 1], pv[(i * 16 + j) * 4]);
      printf("]");
    }
    printf("]}\n");
    free(pc); free(pid); free(pidn); free(pcam); free(pst); free(pm); free(pv);
    pz_filter_destroy(f);
  }

  if (n_devices > 1) {   /* withDevices [0 .. n_devices - 1] */
    check(pz_model_mtt_track((pz_model_desc*)d), "pz_model_mtt_track");
    scenario_lg("multi_mtt_track", d, 256 * 40 * n_devices, n_devices, 1);
    scenario_lg("single_mtt_track_same_n", d, 256 * 40 * n_devices, 1, 1000);
  }
  free(d);
  return 0;
}
