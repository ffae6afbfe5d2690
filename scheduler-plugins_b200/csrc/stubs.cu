// Temporary: plugins not yet built report UNSUPPORTED.
#include "engine.h"
namespace b200s {
int combined_eval(b200s_ctx* c, uint32_t, const int64_t*, int, int) { return c->set_err(B200S_ERR_UNSUPPORTED, "combined not built"); }
}
