// Test infrastructure only (oracle build): a minimal fixed-16-lane stand-in for the subset of
// Google Highway that the reference's hot path touches (src/adaptertrimmer.cpp:61-63,93-96;
// src/sequence.cpp:29-77; src/simdutil.h:9-34).  Written from the op semantics, not from Highway.
// 16 lanes == Highway's static SSE2 target for the reference Makefile's flags (no -march).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>

#define HWY_NAMESPACE fpl_shim16
#define HWY_RESTRICT __restrict__
#define HWY_ATTR
#define HWY_UNLIKELY(x) __builtin_expect(!!(x), 0)
#define HWY_DASSERT(x) ((void)0)
#define HWY_BEFORE_NAMESPACE() static_assert(true, "")
#define HWY_AFTER_NAMESPACE() static_assert(true, "")

namespace hwy {
namespace HWY_NAMESPACE {

constexpr size_t kLanes = 16;

template <typename T> struct ScalableTag { using type = T; };
template <class D> using TFromD = typename D::type;

template <typename T> struct VecN { T v[kLanes]; };
template <typename T> struct MaskN { bool m[kLanes]; };
template <class D> using Vec = VecN<TFromD<D>>;

template <class D> constexpr size_t Lanes(D) { return kLanes; }

template <class D> Vec<D> LoadU(D, const TFromD<D>* p) {
    Vec<D> r; std::memcpy(r.v, p, sizeof(r.v)); return r;
}
// loads n lanes, the remaining lanes are zero
template <class D> Vec<D> LoadN(D, const TFromD<D>* p, size_t n) {
    Vec<D> r; std::memset(r.v, 0, sizeof(r.v));
    if (n > kLanes) n = kLanes;
    std::memcpy(r.v, p, n * sizeof(TFromD<D>)); return r;
}
template <class D> void StoreU(Vec<D> x, D, TFromD<D>* p) { std::memcpy(p, x.v, sizeof(x.v)); }
template <class D> void StoreN(Vec<D> x, D, TFromD<D>* p, size_t n) {
    if (n > kLanes) n = kLanes;
    std::memcpy(p, x.v, n * sizeof(TFromD<D>));
}
template <class D> Vec<D> Set(D, int c) {
    Vec<D> r; for (size_t i = 0; i < kLanes; i++) r.v[i] = (TFromD<D>)c; return r;
}
template <typename T> MaskN<T> Eq(VecN<T> a, VecN<T> b) {
    MaskN<T> r; for (size_t i = 0; i < kLanes; i++) r.m[i] = a.v[i] == b.v[i]; return r;
}
template <typename T> MaskN<T> operator!=(VecN<T> a, VecN<T> b) {
    MaskN<T> r; for (size_t i = 0; i < kLanes; i++) r.m[i] = a.v[i] != b.v[i]; return r;
}
template <typename T> MaskN<T> Or(MaskN<T> a, MaskN<T> b) {
    MaskN<T> r; for (size_t i = 0; i < kLanes; i++) r.m[i] = a.m[i] || b.m[i]; return r;
}
template <typename T> VecN<T> IfThenElse(MaskN<T> m, VecN<T> a, VecN<T> b) {
    VecN<T> r; for (size_t i = 0; i < kLanes; i++) r.v[i] = m.m[i] ? a.v[i] : b.v[i]; return r;
}
template <class D> size_t CountTrue(D, MaskN<TFromD<D>> m) {
    size_t c = 0; for (size_t i = 0; i < kLanes; i++) c += m.m[i] ? 1 : 0; return c;
}
template <class D> Vec<D> Reverse(D, Vec<D> x) {
    Vec<D> r; for (size_t i = 0; i < kLanes; i++) r.v[i] = x.v[kLanes - 1 - i]; return r;
}
// lane i of the result = lane i+amt of the input, upper lanes zero
template <class D> Vec<D> SlideDownLanes(D, Vec<D> x, size_t amt) {
    Vec<D> r; std::memset(r.v, 0, sizeof(r.v));
    for (size_t i = 0; i + amt < kLanes; i++) r.v[i] = x.v[i + amt];
    return r;
}

}  // namespace HWY_NAMESPACE

template <typename T> std::unique_ptr<T[]> AllocateAligned(size_t n) {
    return std::unique_ptr<T[]>(new T[n]);
}
}  // namespace hwy
