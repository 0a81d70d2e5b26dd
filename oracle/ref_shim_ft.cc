// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product.
// Link shim for the full-text merger oracle (oracle/_ref/liboracle_ref_ft.so): the reference's header-only
// ft::Merger (cpp_src/core/ft/ft_fast/mergerimpl.h) + core/ft/idrelset.cc reference a few symbols that live in parts of
// libreindexer we do not build (tools/errors.cc, tools/assertrx.cc, tools/logger.cc, core/ft/config/ftconfig.cc with its JSON
// dependencies, core/rdxcontext.cc, tools/stringstools.cc).  These definitions are ours; they only satisfy the linker.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include "core/ft/config/ftconfig.h"
#include "core/rdxcontext.h"
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

FTConfig::FTConfig(size_t fieldsCount) : fieldsCfg(fieldsCount ? fieldsCount : 1) {}  // real one: core/ft/config/ftconfig.cc
RdxContext::~RdxContext() {}
const char* kDefaultWordPartDelimiters = "-/+_`'";
const char* kDefaultExtraWordsSymbols = "-/+_`'";
void SplitOptions::SetSymbols(std::string_view, std::string_view) {}
}  // namespace reindexer
