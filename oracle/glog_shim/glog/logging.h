// Minimal stand-in for <glog/logging.h>: the Ceres 2.0 public headers only use the CHECK/LOG/VLOG macros
// inside inline code.  Self-contained (no glog symbols).
#pragma once
#include <cstdlib>
#include <iostream>
#include <sstream>
namespace glog_shim {
struct Voidify { void operator&(std::ostream &) {} };
struct Fatal {
    std::ostringstream s;
    Fatal(const char *f, int l) { s << f << ":" << l << " "; }
    ~Fatal() { std::cerr << s.str() << std::endl; std::abort(); }
    std::ostream &stream() { return s; }
};
struct Sink {
    std::ostringstream s;
    std::ostream &stream() { return s; }
};
template <typename T> T &check_notnull(const char *f, int l, const char *n, T &t)
{ if (t == nullptr) { Fatal(f, l).stream() << n; } return t; }
}
#define GLOG_SHIM_FATAL ::glog_shim::Fatal(__FILE__, __LINE__).stream()
#define CHECK(c) (c) ? (void)0 : ::glog_shim::Voidify() & GLOG_SHIM_FATAL << "Check failed: " #c " "
#define CHECK_OP_(a, b, op) CHECK((a) op (b))
#define CHECK_EQ(a, b) CHECK_OP_(a, b, ==)
#define CHECK_NE(a, b) CHECK_OP_(a, b, !=)
#define CHECK_LE(a, b) CHECK_OP_(a, b, <=)
#define CHECK_LT(a, b) CHECK_OP_(a, b, <)
#define CHECK_GE(a, b) CHECK_OP_(a, b, >=)
#define CHECK_GT(a, b) CHECK_OP_(a, b, >)
#define CHECK_NOTNULL(p) ::glog_shim::check_notnull(__FILE__, __LINE__, "'" #p "' must be non NULL", (p))
#define DCHECK(c) while (false) CHECK(c)
#define DCHECK_EQ(a, b) while (false) CHECK_EQ(a, b)
#define DCHECK_NE(a, b) while (false) CHECK_NE(a, b)
#define DCHECK_LE(a, b) while (false) CHECK_LE(a, b)
#define DCHECK_LT(a, b) while (false) CHECK_LT(a, b)
#define DCHECK_GE(a, b) while (false) CHECK_GE(a, b)
#define DCHECK_GT(a, b) while (false) CHECK_GT(a, b)
#define GLOG_SHIM_SEV_INFO ::glog_shim::Sink().stream()
#define GLOG_SHIM_SEV_WARNING ::glog_shim::Sink().stream()
#define GLOG_SHIM_SEV_ERROR ::glog_shim::Sink().stream()
#define GLOG_SHIM_SEV_FATAL GLOG_SHIM_FATAL
#define LOG(sev) GLOG_SHIM_SEV_##sev
#define LOG_IF(sev, c) !(c) ? (void)0 : ::glog_shim::Voidify() & LOG(sev)
#define VLOG(n) true ? (void)0 : ::glog_shim::Voidify() & ::glog_shim::Sink().stream()
#define VLOG_IF(n, c) true ? (void)0 : ::glog_shim::Voidify() & ::glog_shim::Sink().stream()
#define VLOG_IS_ON(n) false
