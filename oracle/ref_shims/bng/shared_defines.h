/* placeholder for libbng's shared_defines.h (absent sibling repo); src/util.c needs nothing from it
 * for the functions oracle/_ref exposes.  TEST INFRASTRUCTURE, own file. */
#pragma once
