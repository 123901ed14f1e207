// Library-level C-ABI entry points: version and thread-local error text.
#include "ugl_host.cuh"

namespace ugl {
char* last_error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}
}  // namespace ugl

extern "C" int ugl_version(void) { return UGL_VERSION; }
extern "C" const char* ugl_last_error(void) { return ugl::last_error_buffer(); }
