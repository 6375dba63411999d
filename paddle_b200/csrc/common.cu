#include "include/b200_common.cuh"
#include "include/b200_ops.h"
#include <stdio.h>
#include <string.h>

namespace b200 {
static thread_local char g_err[512] = {0};
static thread_local char g_err_out[512] = {0};
void set_last_error(const char* file, int line, const char* msg) {
  if (g_err[0] == 0) snprintf(g_err, sizeof(g_err), "%s:%d: %s", file, line, msg);
}
const char* take_last_error() {
  memcpy(g_err_out, g_err, sizeof(g_err));
  g_err[0] = 0;
  return g_err_out;
}
}  // namespace b200
