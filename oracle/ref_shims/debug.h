// Placeholder for include/debug.h of the reference (oracle/_ref build only): that header pulls MCell4's
// src4/defines.h and through it the absent libbng; the MCell3 translation units compiled here only use its
// dump helpers inside DEBUG_* blocks that are disabled in release builds (include/debug_config.h).
#pragma once
#include <iostream>
#include <string>
#include <vector>
