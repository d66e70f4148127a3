// L1-ball QP solver (placeholder translation unit; the solver lands in a later commit).
#include "kf_internal.h"
