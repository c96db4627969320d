/* oracle/ref_shims/config.h — build-time shim (TEST INFRASTRUCTURE, own file): the reference's autoconf step
 * would copy src/config-nix.h to config.h; include it where it lies instead. */
#include "config-nix.h"
