// Shim: absl::flat_hash_set -> std::unordered_set (included by the reference, never used).
#pragma once
#include <unordered_set>
#include "flat_hash_map.h"
namespace absl {
template <class K>
using flat_hash_set = std::unordered_set<K, shim_detail::Hash<K>>;
}
