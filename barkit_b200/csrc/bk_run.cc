// placeholder until the file->file runtime lands
#include "../../include/barkit_b200.h"
#include <string.h>
extern "C" int bk_run(const bk_run_args* args, char* err, size_t errcap) {
  (void)args;
  if (err && errcap) strncpy(err, "bk_run not built yet", errcap - 1), err[errcap - 1] = 0;
  return BK_E_UNSUPPORTED;
}
