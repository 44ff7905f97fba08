/* oracle/shim/utils/flog.h -- TEST INFRASTRUCTURE: the reference's logger (core/src/utils/flog.h), silent, so that
 * decoder_modules/radio/src/rds.cpp compiles from where it lies. */
#pragma once
namespace flog {
    template <class... A> inline void debug(const char*, A...) {}
    template <class... A> inline void info(const char*, A...) {}
    template <class... A> inline void warn(const char*, A...) {}
    template <class... A> inline void error(const char*, A...) {}
}
