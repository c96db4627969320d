/* oracle/ref_shims4/bng/shared_defines.h — TEST INFRASTRUCTURE, own file (oracle/_ref build only).
 * Stand-in for libbng's bng/shared_defines.h (sibling repo mcellteam/libbng, absent and unpinned: SURVEY 8c), just
 * enough for the reference's src4/defines.h to compile so that src4/collision_utils_subparts.inl can be built
 * UNMODIFIED into oracle/_ref/libmcell4ref.so.  Values restated from the MCell3 originals in the reference tree:
 * EPS_C = 1e-12, SQRT_EPS_C = 1e-6 (src/mcell_structs.h:238-239), distinguishable (src/util.c:449-463);
 * cmp_eq(a, b, eps) := fabs(a - b) < eps is an assumption.  INDEXER_WA selects the std containers of
 * defines.h:309-319 (no boost). */
#pragma once
#include <cmath>
#include <cfloat>
#include <cstdint>
#include <set>
#include <string>
#include <vector>
#define INDEXER_WA
/* double-precision positions: the production configuration (every POS_T_BYTES == 4 block is the
 * experimental float build; defines.h:352-359 picks glm::dvec3 for 8) */
#define POS_T_BYTES 8
#define FLOAT_T_BYTES 8
typedef unsigned int uint;
namespace BNGCommon {
typedef double pos_t; typedef double stime_t;
const double EPS = 1e-12, SQRT_EPS = 1e-6, DBL_GIGANTIC = 1e140;
const pos_t POS_EPS = 1e-12, POS_SQRT_EPS = 1e-6, POS_GIGANTIC = 1e140;
const stime_t STIME_EPS = 1e-12, STIME_SQRT_EPS = 1e-6, STIME_GIGANTIC = 1e140;
static inline double fabs_f(double x) { return std::fabs(x); }
static inline bool cmp_eq(double a, double b, double eps = EPS) { return std::fabs(a - b) < eps; }
static inline bool distinguishable_f(double a, double b, double eps) {
  double c = std::fabs(a - b); a = std::fabs(a); if (a < 1) a = 1; b = std::fabs(b);
  if (b < a) eps *= a; else eps *= b; return c > eps; }
static inline bool distinguishable_p(double a, double b, double eps) { return distinguishable_f(a, b, eps); }
static inline double sqrt_f(double x) { return std::sqrt(x); }
static inline double pow_f(double a, double b) { return std::pow(a, b); }
static inline double floor_f(double x) { return std::floor(x); }
static inline double round_f(double x) { return std::round(x); }
static inline std::string f_to_str(double v, int = 8) { return std::to_string(v); }
}
namespace BNG {
typedef uint species_id_t; const species_id_t SPECIES_ID_INVALID = 0xFFFFFFFFu;
typedef int orientation_t;
const orientation_t ORIENTATION_DOWN = -1, ORIENTATION_NONE = 0, ORIENTATION_UP = 1, ORIENTATION_NOT_SET = 2, ORIENTATION_DEPENDS_ON_SURF_COMP = 3;
}
const uint ID_INVALID = 0xFFFFFFFFu, ID_INVALID2 = 0xFFFFFFFEu, INDEX_INVALID = 0xFFFFFFFFu, INDEX_INVALID2 = 0xFFFFFFFEu;
template <typename T> class uint_set : public std::set<T> {
public:
  void insert_unique(const T v) { this->insert(v); }
  void erase_existing(const T v) { this->erase(v); }
};
