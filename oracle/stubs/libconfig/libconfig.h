/* oracle/stubs/libconfig/libconfig.h -- TEST INFRASTRUCTURE.  libconfig is not in this
 * image; these declarations let `gcc -fsyntax-only` parse the (patched) reference
 * src/config.c so that integration/cuda_cf32.patch is at least syntax-checked there
 * (oracle/build_ref.sh).  Nothing is linked against this. */
#ifndef XL_STUB_LIBCONFIG_H
#define XL_STUB_LIBCONFIG_H
#define CONFIG_TRUE 1
#define CONFIG_FALSE 0
typedef struct config_setting_t { int type; } config_setting_t;
typedef struct config_t { int dummy; } config_t;
void config_init(config_t *config);
void config_destroy(config_t *config);
int config_read_file(config_t *config, const char *filename);
config_setting_t *config_lookup(const config_t *config, const char *path);
int config_setting_get_int(const config_setting_t *setting);
double config_setting_get_float(const config_setting_t *setting);
int config_setting_get_bool(const config_setting_t *setting);
const char *config_setting_get_string(const config_setting_t *setting);
const char *config_error_text(const config_t *config);
const char *config_error_file(const config_t *config);
int config_error_line(const config_t *config);
#endif
