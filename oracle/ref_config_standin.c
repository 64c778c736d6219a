/*
 * oracle/ref_config_standin.c -- TEST INFRASTRUCTURE.  A stand-in for the reference's
 * src/config.c, which cannot be built here (it parses with libconfig, absent from this
 * image), so that the reference's own server test test/test_tcp_server.c can be compiled
 * UNMODIFIED with its real tcp_server.c / dsp_worker.c / queue.c / sdr_device.c
 * (oracle/Makefile: _ref/test_tcp_server_*).  It implements the same two entry points
 * (src/config.h: create_server_config, destroy_server_config) for the flat
 * `key=value` files the reference's tests use, with the defaults of src/config.c:98-264.
 *
 * XL_TEST_CPU_OPTIMIZATION=<NATIVE_CF32|OPTIMIZED_CF32|CUDA_CF32> overrides the
 * file's cpu_optimization, so the same unmodified test program can be pointed at each
 * mode (CUDA_CF32 exists only in the tree patched with integration/cuda_cf32.patch: this
 * file is compiled against whichever config.h the build uses).
 */
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "config.h"

static char *dup_str(const char *s) {
  char *r = (char *)malloc(strlen(s) + 1);
  if (r != NULL) strcpy(r, s);
  return r;
}

static const char *lookup(char keys[][64], char values[][256], int n, const char *key) {
  for (int i = 0; i < n; i++)
    if (strcmp(keys[i], key) == 0) return values[i];
  return NULL;
}

static int as_int(const char *v, int fallback) { return v != NULL ? (int)strtod(v, NULL) : fallback; }

static int as_bool(const char *v, int fallback) {
  if (v == NULL) return fallback;
  return strcmp(v, "true") == 0 || strcmp(v, "1") == 0;
}

static int parse_optimization(const char *s, cpu_optimization *out) {
  if (strcmp(s, "NATIVE_CF32") == 0) {
    *out = NATIVE_CF32;
    return 0;
  }
  if (strcmp(s, "OPTIMIZED_CF32") == 0) {
    *out = OPTIMIZED_CF32;
    return 0;
  }
#ifdef XL_HAVE_CUDA_CF32
  if (strcmp(s, "CUDA_CF32") == 0) {
    *out = CUDA_CF32;
    return 0;
  }
#endif
  return -1;
}

int create_server_config(struct server_config **config, const char *path) {
  fprintf(stdout, "loading configuration from: %s\n", path);
  FILE *f = fopen(path, "r");
  if (f == NULL) {
    fprintf(stderr, "<3>unable to read configuration: %s\n", path);
    return -1;
  }
  static char keys[64][64], values[64][256];
  int n = 0;
  char line[512];
  while (n < 64 && fgets(line, sizeof(line), f) != NULL) {
    char *hash = strchr(line, '#');
    if (hash != NULL) *hash = '\0';
    char *eq = strchr(line, '=');
    if (eq == NULL) continue;
    *eq = '\0';
    char *k = line, *v = eq + 1;
    while (*k == ' ' || *k == '\t') k++;
    char *ke = k + strlen(k);
    while (ke > k && (ke[-1] == ' ' || ke[-1] == '\t')) *--ke = '\0';
    while (*v == ' ' || *v == '\t' || *v == '"') v++;
    char *ve = v + strlen(v);
    while (ve > v && (ve[-1] == '\n' || ve[-1] == '\r' || ve[-1] == ' ' || ve[-1] == ';' || ve[-1] == '"')) *--ve = '\0';
    snprintf(keys[n], sizeof(keys[n]), "%s", k);
    snprintf(values[n], sizeof(values[n]), "%s", v);
    n++;
  }
  fclose(f);
#define GET(key) lookup(keys, values, n, key)
  struct server_config *r = (struct server_config *)calloc(1, sizeof(*r));
  if (r == NULL) return -ENOMEM;
  r->sdr_type = (sdr_type_t)as_int(GET("sdr_type"), 0);
  r->gain_mode = as_int(GET("gain_mode"), 0);
  r->gain = (int)(strtod(GET("gain") != NULL ? GET("gain") : "0", NULL) * 10);
  r->bias_t = as_int(GET("bias_t"), 0);
  r->ppm = as_int(GET("ppm"), 0);
  r->device_index = as_int(GET("device_index"), 0);
  r->band_sampling_rate = (uint32_t)as_int(GET("band_sampling_rate"), 2400000);
  r->queue_size = as_int(GET("queue_size"), 64);
  if (r->queue_size <= 0) {
    fprintf(stderr, "<3>queue size should be positive: %d\n", r->queue_size);
    free(r);
    return -1;
  }
  r->buffer_size = (uint32_t)as_int(GET("buffer_size"), 262144);
  if (r->sdr_type == SDR_TYPE_AIRSPY && r->buffer_size != 262144) r->buffer_size = 262144;
  r->lpf_cutoff_rate = as_int(GET("lpf_cutoff_rate"), 5);
  r->bind_address = dup_str(GET("bind_address") != NULL ? GET("bind_address") : "127.0.0.1");
  r->port = as_int(GET("port"), 8090);
  r->read_timeout_seconds = as_int(GET("read_timeout_seconds"), 5);
  if (r->read_timeout_seconds <= 0) {
    fprintf(stderr, "<3>read timeout should be positive: %d\n", r->read_timeout_seconds);
    destroy_server_config(r);
    return -1;
  }
  const char *default_folder = getenv("TMPDIR");
  if (default_folder == NULL) default_folder = "/tmp";
  r->base_path = dup_str(GET("base_path") != NULL ? GET("base_path") : default_folder);
  r->use_gzip = as_bool(GET("use_gzip"), 1);
  r->airspy_gain_mode = (airspy_gain_mode_t)as_int(GET("airspy_gain_mode"), 0);
  r->airspy_vga_gain = as_int(GET("airspy_vga_gain"), 5);
  r->airspy_mixer_gain = as_int(GET("airspy_mixer_gain"), 13);
  r->airspy_lna_gain = as_int(GET("airspy_lna_gain"), 14);
  r->airspy_linearity_gain = as_int(GET("airspy_linearity_gain"), 0);
  r->airspy_sensitivity_gain = as_int(GET("airspy_sensitivity_gain"), 0);
  r->hackrf_bias_t = (uint8_t)as_int(GET("hackrf_bias_t"), 0);
  r->hackrf_amp = as_int(GET("hackrf_amp"), 0);
  r->hackrf_lna_gain = as_int(GET("hackrf_lna_gain"), 16);
  r->hackrf_vga_gain = as_int(GET("hackrf_vga_gain"), 16);
  r->optimization = NATIVE_CF32;
  const char *opt = getenv("XL_TEST_CPU_OPTIMIZATION");
  if (opt == NULL) opt = GET("cpu_optimization");
  if (opt != NULL && parse_optimization(opt, &r->optimization) != 0) {
    fprintf(stderr, "<3>invalid cpu_optimization: %s\n", opt);
    destroy_server_config(r);
    return -1;
  }
  fprintf(stdout, "cpu_optimization: %d\n", (int)r->optimization);
#undef GET
  *config = r;
  return 0;
}

void destroy_server_config(struct server_config *config) {
  if (config == NULL) return;
  free(config->bind_address);
  free(config->base_path);
  free(config->device_serial);
  free(config);
}
