// oracle/shim/spdlog/spdlog.h — TEST INFRASTRUCTURE ONLY.
//
// The reference pins spdlog v1.17.0 (third_party/CMakeLists.txt:62-68), fetched
// at configure time. Its hot path only calls spdlog::should_log(level) and
// spdlog::{error,warn,info,debug,trace,critical}("{}", msg)
// (utils/structured_log.h:35,223-269). Logging is not part of any result, so
// this stub reports every level as disabled and drops the message.
#pragma once

namespace spdlog {
namespace level {
enum level_enum : int { trace = 0, debug = 1, info = 2, warn = 3, err = 4, critical = 5, off = 6 };
}  // namespace level

inline bool should_log(level::level_enum /*lvl*/) { return false; }

template <typename... Args>
inline void trace(const Args&... /*args*/) {}
template <typename... Args>
inline void debug(const Args&... /*args*/) {}
template <typename... Args>
inline void info(const Args&... /*args*/) {}
template <typename... Args>
inline void warn(const Args&... /*args*/) {}
template <typename... Args>
inline void error(const Args&... /*args*/) {}
template <typename... Args>
inline void critical(const Args&... /*args*/) {}
}  // namespace spdlog
