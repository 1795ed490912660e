// oracle/shim/absl/container/flat_hash_set.h — TEST INFRASTRUCTURE ONLY (see flat_hash_map.h).
#pragma once

#include <functional>
#include <unordered_set>

namespace absl {

template <typename K, typename Hash = std::hash<K>, typename Eq = std::equal_to<K>>
class flat_hash_set : public std::unordered_set<K, Hash, Eq> {
 public:
  using Base = std::unordered_set<K, Hash, Eq>;
  using Base::Base;
  [[nodiscard]] size_t capacity() const { return this->bucket_count(); }
};

}  // namespace absl
