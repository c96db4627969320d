/* oracle/ref_shims/nfsim_c.h — placeholder for the absent NFsim C interface (sibling repo nfsimCInterface,
 * CMakeLists.txt:146-148).  Only opaque value types are needed to parse src/react.h; no NFsim code path is
 * reachable from the functions oracle/_ref exposes.  TEST INFRASTRUCTURE, own file. */
#pragma once
typedef struct queryOptions_ { int unused; } queryOptions;
typedef struct queryResults_ { int unused; } queryResults;
typedef struct reactantQueryResults_ { int unused; } reactantQueryResults;
typedef struct reactionResult_ { int unused; } reactionResult;
