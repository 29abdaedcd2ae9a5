// apk_ref_stub.hpp -- ORACLE BUILD SUPPORT (test infrastructure, NOT product code).
//
// Minimal stand-in for the slice of the Parthenon/Kokkos API that AthenaPK's hot-path
// *headers* (src/recon/*_simple.hpp, src/hydro/rsolvers/*.hpp, src/eos/adiabatic_*.hpp)
// touch, so that those headers can be compiled UNMODIFIED, from where they lie under
// /root/reference, with plain g++ (no cmake, no Kokkos, no generated config.hpp).
// Serial semantics only: par_for_inner is a for loop, scratch pads are heap arrays.
// Mirrors (names/semantics only, code is our own):
//   external/parthenon/src/kokkos_abstraction.hpp:134-139,924-991 (team_mbr_t, ScratchPad2D,
//   par_for_inner), interface/variable_pack.hpp:258-401 (VariablePack / VariableFluxPack),
//   config.hpp.in:86-89 (TINY_NUMBER, SQR, SIGN), basic_types.hpp:36 (Real = double).
#ifndef APK_REF_STUB_HPP_
#define APK_REF_STUB_HPP_

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <limits>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <type_traits>
#include <vector>

#define KOKKOS_INLINE_FUNCTION inline
#define KOKKOS_FORCEINLINE_FUNCTION inline
#define KOKKOS_LAMBDA [=]
#define TINY_NUMBER 1.0e-20
#define SQR(x) ((x) * (x))
#define SIGN(x) (((x) < 0.0) ? -1.0 : 1.0)
#define PARTHENON_FAIL(msg)                                                              \
  {                                                                                      \
    std::fprintf(stderr, "PARTHENON_FAIL\n");                                            \
    std::abort();                                                                        \
  }
// A failed requirement is reported to the shim (so tests can observe it) rather than abort.
extern "C" int apk_ref_require_failed;
#define PARTHENON_REQUIRE(cond, msg)                                                     \
  if (!(cond)) apk_ref_require_failed = 1;
#define PARTHENON_REQUIRE_THROWS(cond, msg) PARTHENON_REQUIRE(cond, msg)

namespace parthenon {
using Real = double;
constexpr int X1DIR = 1, X2DIR = 2, X3DIR = 3;
enum CoordinateDirection { NODIR = 0 };

struct team_mbr_t {};

template <class F>
inline void par_for_inner(team_mbr_t const &, const int il, const int iu, const F &f) {
  for (int i = il; i <= iu; ++i) f(i);
}

template <typename T>
class ScratchPad2D {
 public:
  ScratchPad2D(int n, int ni) : ni_(ni), store_(std::make_shared<std::vector<T>>(n * ni)) {
    p_ = store_->data();
  }
  T &operator()(int n, int i) const { return p_[n * ni_ + i]; }
  T *data() const { return p_; }
  void assign_data(T *p) { p_ = p; }

 private:
  int ni_;
  std::shared_ptr<std::vector<T>> store_;
  T *p_;
};

struct StubCoords {
  Real dx[3] = {1.0, 1.0, 1.0};
  template <int dir>
  Real Dxc(int, int, int) const {
    return dx[dir - 1];
  }
};

template <typename T>
class VariablePack {
 public:
  VariablePack() = default;
  VariablePack(int nv, int nk, int nj, int ni)
      : nv_(nv), nk_(nk), nj_(nj), ni_(ni),
        store_(std::make_shared<std::vector<T>>((size_t)nv * nk * nj * ni)) {}
  T &operator()(int n, int k, int j, int i) const {
    return (*store_)[(((size_t)n * nk_ + k) * nj_ + j) * ni_ + i];
  }
  int GetDim(int d) const { return d == 4 ? nv_ : d == 3 ? nk_ : d == 2 ? nj_ : ni_; }
  const StubCoords &GetCoords() const { return coords; }
  StubCoords coords;

 protected:
  int nv_ = 0, nk_ = 0, nj_ = 0, ni_ = 0;
  std::shared_ptr<std::vector<T>> store_;
};

template <typename T>
class VariableFluxPack : public VariablePack<T> {
 public:
  VariableFluxPack() = default;
  VariableFluxPack(int nv, int nk, int nj, int ni)
      : VariablePack<T>(nv, nk, nj, ni),
        flux_(std::make_shared<std::vector<T>>((size_t)3 * nv * nk * nj * ni)) {}
  T &flux(int dir, int n, int k, int j, int i) const {
    return (*flux_)[((((size_t)(dir - 1) * this->nv_ + n) * this->nk_ + k) * this->nj_ + j) *
                        this->ni_ + i];
  }

 private:
  std::shared_ptr<std::vector<T>> flux_;
};

// ---- slice used by prolong_restrict/pr_ops.hpp and src/hydro/prolongation/custom_ops.hpp ----
// (basic_types.hpp:159-200 TopologicalElement values/offsets, mesh/domain.hpp IndexRange,
//  coordinates/uniform_coordinates.hpp:31-58,118-121,168-181,284-307 -- restated, serial)
enum class TopologicalElement : std::size_t { CC = 0, F1 = 3, F2 = 4, F3 = 5, E1 = 6, E2 = 7, E3 = 8, NN = 9 };
using TE = TopologicalElement;
inline constexpr int TopologicalOffsetI(TE el) { return (el == TE::F1 || el == TE::E2 || el == TE::E3 || el == TE::NN); }
inline constexpr int TopologicalOffsetJ(TE el) { return (el == TE::F2 || el == TE::E3 || el == TE::E1 || el == TE::NN); }
inline constexpr int TopologicalOffsetK(TE el) { return (el == TE::F3 || el == TE::E2 || el == TE::E1 || el == TE::NN); }
enum class TopologicalType { Cell, Face, Edge, Node };
inline constexpr TopologicalType GetTopologicalType(TE el) {
  return el == TE::CC ? TopologicalType::Cell
                      : el == TE::NN ? TopologicalType::Node
                                     : (el == TE::F1 || el == TE::F2 || el == TE::F3) ? TopologicalType::Face
                                                                                      : TopologicalType::Edge;
}
// only needed so that pr_ops.hpp's unused internal-prolongation operators parse
inline constexpr bool IsSubmanifold(TE a, TE b) { return a == b; }
struct IndexRange {
  int s = 0, e = 0;
};
struct VariableState {};
struct Coordinates_t {
  Real dx_[3] = {1, 1, 1}, xmin_[3] = {0, 0, 0}, area_[3] = {1, 1, 1}, cell_volume_ = 1;
  int istart_[3] = {0, 0, 0};
  Coordinates_t() = default;
  // uniform_coordinates.hpp:31-42 (xmin_ = block xmin - istart*dx)
  Coordinates_t(const Real xmin_block[3], const Real dx[3], const int istart[3]) {
    for (int d = 0; d < 3; ++d) {
      dx_[d] = dx[d];
      istart_[d] = istart[d];
      xmin_[d] = xmin_block[d] - istart_[d] * dx_[d];
    }
    finish();
  }
  // uniform_coordinates.hpp:43-58
  Coordinates_t(const Coordinates_t &src, int coarsen) {
    for (int d = 0; d < 3; ++d) {
      istart_[d] = src.istart_[d];
      dx_[d] = src.dx_[d];
      xmin_[d] = src.xmin_[d];
    }
    xmin_[0] += istart_[0] * dx_[0] * (1 - coarsen);
    xmin_[1] += istart_[1] * dx_[1] * (1 - coarsen);
    xmin_[2] += istart_[2] * dx_[2] * (1 - coarsen);
    dx_[0] *= coarsen;
    dx_[1] *= (istart_[1] > 0 ? coarsen : 1);
    dx_[2] *= (istart_[2] > 0 ? coarsen : 1);
    finish();
  }
  void finish() {
    area_[0] = dx_[1] * dx_[2];
    area_[1] = dx_[0] * dx_[2];
    area_[2] = dx_[0] * dx_[1];
    cell_volume_ = dx_[0] * dx_[1] * dx_[2];
  }
  Real Dxc(const int dir, int, int, int) const { return dx_[dir - 1]; }
  template <int dir>
  Real Xc(const int idx) const {
    return xmin_[dir - 1] + (idx + 0.5) * dx_[dir - 1];
  }
  template <int dir, TopologicalElement el>
  Real X(const int idx) const {
    if constexpr (dir == 1 && TopologicalOffsetI(el)) return xmin_[dir - 1] + idx * dx_[dir - 1];
    else if constexpr (dir == 2 && TopologicalOffsetJ(el)) return xmin_[dir - 1] + idx * dx_[dir - 1];
    else if constexpr (dir == 3 && TopologicalOffsetK(el)) return xmin_[dir - 1] + idx * dx_[dir - 1];
    else return Xc<dir>(idx);
  }
  template <TopologicalElement el>
  Real Volume(const int, const int, const int) const {
    if constexpr (el == TE::CC) return cell_volume_;
    else if constexpr (el == TE::F1) return area_[0];
    else if constexpr (el == TE::F2) return area_[1];
    else if constexpr (el == TE::F3) return area_[2];
    else if constexpr (el == TE::E1) return dx_[0];
    else if constexpr (el == TE::E2) return dx_[1];
    else if constexpr (el == TE::E3) return dx_[2];
    else return 1.0;
  }
};
// 7-index accessor (element, l, m, n, k, j, i) over a caller-owned [nv][nk][nj][ni] array
template <typename T, typename State = VariableState>
class ParArrayND {
 public:
  ParArrayND() = default;
  ParArrayND(T *p, int nv, int nk, int nj, int ni) : p_(p), nv_(nv), nk_(nk), nj_(nj), ni_(ni) {}
  T &operator()(int, int, int, int n, int k, int j, int i) const {
    return p_[(((size_t)n * nk_ + k) * nj_ + j) * ni_ + i];
  }

 private:
  T *p_ = nullptr;
  int nv_ = 0, nk_ = 0, nj_ = 0, ni_ = 0;
};
template <typename T, typename State = VariableState>
using ParArray6D = ParArrayND<T, State>;

template <typename T> class ParArray4D {};
template <typename T> class MeshData {};
template <typename T> class MeshBlockData {};
template <typename T> class MeshBlockVarPack {};
class MeshBlock {};
class Mesh {};
class ParameterInput {};
class StateDescriptor {};
struct SimTime {};
using Packages_t = int;
enum class TaskStatus { complete, incomplete, iterate, fail };
enum class AmrTag : int { derefine = -1, same = 0, refine = 1 };
namespace package { namespace prelude {
using parthenon::AmrTag;
using parthenon::Mesh;
using parthenon::MeshBlockData;
using parthenon::MeshData;
using parthenon::ParameterInput;
using parthenon::Real;
using parthenon::StateDescriptor;
using parthenon::TaskStatus;
} }
} // namespace parthenon

using parthenon::IndexRange;
using parthenon::ParArrayND;
using parthenon::Real;
using parthenon::VariableFluxPack;
using parthenon::VariablePack;

#endif // APK_REF_STUB_HPP_
