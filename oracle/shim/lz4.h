// oracle/shim/lz4.h — TEST INFRASTRUCTURE ONLY.
//
// The reference's query-result cache compresses entries with LZ4
// (src/cache/result_compressor.cpp:39-91; system liblz4, absent here). The cache
// is disabled on the measured path (cache_manager == nullptr, as in
// docs/releases/v1.3.5.md:208), so these "stored block" stand-ins only exist to
// let src/cache/*.cpp link; they are never executed by the oracle.
#pragma once
#include <cstring>

inline int LZ4_compressBound(int input_size) { return input_size < 0 ? 0 : input_size + 16; }
inline int LZ4_compress_default(const char* src, char* dst, int src_size, int dst_capacity) {
  if (src_size < 0 || dst_capacity < src_size) {
    return 0;
  }
  std::memcpy(dst, src, static_cast<size_t>(src_size));
  return src_size;
}
inline int LZ4_decompress_safe(const char* src, char* dst, int compressed_size, int dst_capacity) {
  if (compressed_size < 0 || dst_capacity < compressed_size) {
    return -1;
  }
  std::memcpy(dst, src, static_cast<size_t>(compressed_size));
  return compressed_size;
}
