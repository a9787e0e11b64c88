// Shim: absl::flat_hash_map -> std::unordered_map with hashes for the key
// types the reference uses (uint64, unsigned __int128, pair<int,uint128>).
#pragma once
#include <unordered_map>
#include <utility>
#include <cstdint>
namespace absl {
namespace shim_detail {
inline uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
template <class K> struct Hash { size_t operator()(const K& k) const { return std::hash<K>()(k); } };
template <> struct Hash<uint64_t> { size_t operator()(uint64_t k) const { return (size_t)mix64(k); } };
template <> struct Hash<unsigned __int128> {
    size_t operator()(unsigned __int128 k) const {
        return (size_t)mix64((uint64_t)k ^ mix64((uint64_t)(k >> 64) + 0x9e3779b97f4a7c15ULL));
    }
};
template <class A, class B> struct Hash<std::pair<A, B>> {
    size_t operator()(const std::pair<A, B>& p) const {
        return (size_t)mix64((uint64_t)Hash<A>()(p.first) * 0x9e3779b97f4a7c15ULL + (uint64_t)Hash<B>()(p.second));
    }
};
template <> struct Hash<int> { size_t operator()(int k) const { return (size_t)mix64((uint64_t)(uint32_t)k); } };
}  // namespace shim_detail
template <class K, class V>
using flat_hash_map = std::unordered_map<K, V, shim_detail::Hash<K>>;
}  // namespace absl
