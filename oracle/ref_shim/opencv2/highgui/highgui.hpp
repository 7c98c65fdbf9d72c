/* header shim: routes the reference include to oracle/ref_shim/cvshim.h (test infrastructure) */
#include "cvshim.h"
