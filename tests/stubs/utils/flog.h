// tests/stubs/utils/flog.h -- TEST INFRASTRUCTURE: the reference's logger, silent
#pragma once
namespace flog {
    template <class... A> inline void debug(const char*, A...) {}
    template <class... A> inline void info(const char*, A...) {}
    template <class... A> inline void warn(const char*, A...) {}
    template <class... A> inline void error(const char*, A...) {}
}
