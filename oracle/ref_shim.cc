// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product (librxgpu.so).
//
// Link shim for compiling a handful of reference translation units *in place* from
// /root/reference/cpp_src (never copied) into oracle/_ref/.  The reference TUs
// reference a few symbols that live in parts of libreindexer we do not build
// (tools/errors.cc, tools/assertrx.cc, tools/logger.cc, core/definitions/
// quantization_config.cc).  These definitions are ours; they only satisfy the linker.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include "core/definitions/quantization_config.h"
#include "core/index/float_vector/hnswlib/hnsw_interface.h"
#include "tools/errors.h"

namespace reindexer {
const Error::WhatPtr Error::defaultErrorText_{make_intrusive<Error::WhatT>("Error text generation failed")};
void fail_assertrx(const char* assertion, const char* file, unsigned line, const char* func) noexcept {
	std::fprintf(stderr, "oracle/_ref: assertion failed: %s (%s:%u %s)\n", assertion, file, line, func);
	std::abort();
}
void fail_throwrx(const char* assertion, const char*, unsigned, const char*) noexcept(false) { throw std::runtime_error(assertion); }
namespace logger_details {
std::atomic<int> g_LogLevel{0};
void logPrintImpl(int, char* buf) { std::fputs(buf, stderr); }
}  // namespace logger_details
}  // namespace reindexer

namespace hnswlib {
void QuantizationConfig::Deserialize(IReader&) {}
void QuantizationConfig::Serialize(IWriter&) const {}
}  // namespace hnswlib
