// oracle/shim/absl/container/flat_hash_map.h — TEST INFRASTRUCTURE ONLY.
//
// The reference pins abseil 20250127.1 (third_party/CMakeLists.txt:172-181) and
// fetches it at configure time; it is not in this offline image. The hot-path
// sources only use absl::flat_hash_map as an unordered associative container
// (plus `capacity()` for memory accounting, storage/document_store.cpp:486-530),
// so std::unordered_map is a drop-in. Iteration order is unspecified in both
// and is never observable in search results. C++20 is required so that
// find(std::string_view) works through the transparent hash/equal functors
// (utils/hash_utils.h).
#pragma once

#include <functional>
#include <unordered_map>

namespace absl {

template <typename K, typename V, typename Hash = std::hash<K>, typename Eq = std::equal_to<K>>
class flat_hash_map : public std::unordered_map<K, V, Hash, Eq> {
 public:
  using Base = std::unordered_map<K, V, Hash, Eq>;
  using Base::Base;
  [[nodiscard]] size_t capacity() const { return this->bucket_count(); }
};

}  // namespace absl
