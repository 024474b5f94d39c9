// oracle/ref_cec_tables.cpp - TEST INFRASTRUCTURE ONLY.
// Supplies the data symbols the unmodified reference cec2014.cpp / cec2013.cpp link against
// (declared `extern const` in src/problems/cec2014_data.hpp:46-48 and cec2013_data.hpp:46-47; their
// real definitions are in .MISSING_LARGE_BLOBS).  They are filled on demand from oracle/cec_synth.c.
// This TU deliberately does not include the reference's *_data.hpp: the objects are defined
// non-const here so that ref_capi.cpp can populate them lazily, before a cec UDP is constructed.
#include <mutex>
#include <unordered_map>
#include <vector>

#include "cec_synth.h"

namespace pagmo { namespace detail {
namespace cec2014_data {
std::unordered_map<unsigned, std::unordered_map<unsigned, std::vector<double>>> rotation_data;
std::unordered_map<unsigned, std::unordered_map<unsigned, std::vector<int>>> shuffle_data;
std::unordered_map<unsigned, std::vector<double>> shift_data;
}
namespace cec2013_data {
std::unordered_map<unsigned, std::vector<double>> MD;
std::vector<double> shift_data;
}
}}

namespace oracle_ref {

static std::mutex g_tables_mutex;

void ensure_cec2014_tables(unsigned func, unsigned dim)
{
    namespace d = pagmo::detail::cec2014_data;
    std::lock_guard<std::mutex> lk(g_tables_mutex);
    auto &rot = d::rotation_data[func];
    if (!rot.count(dim)) {
        std::vector<double> m(static_cast<std::size_t>(CEC_SYNTH_NCOMP) * dim * dim);
        cec2014_synth_rotation(func, dim, m.data());
        rot.emplace(dim, std::move(m));
    }
    if (!d::shift_data.count(func)) {
        std::vector<double> s(static_cast<std::size_t>(CEC_SYNTH_NCOMP) * 100);
        cec2014_synth_shift(func, s.data());
        d::shift_data.emplace(func, std::move(s));
    }
    auto &shuf = d::shuffle_data[func];
    if (!shuf.count(dim)) {
        std::vector<int> p(static_cast<std::size_t>(CEC_SYNTH_NCOMP) * dim);
        cec2014_synth_shuffle(func, dim, p.data());
        shuf.emplace(dim, std::move(p));
    }
}

void ensure_cec2013_tables(unsigned dim)
{
    namespace d = pagmo::detail::cec2013_data;
    std::lock_guard<std::mutex> lk(g_tables_mutex);
    if (!d::MD.count(dim)) {
        std::vector<double> m(static_cast<std::size_t>(CEC_SYNTH_NCOMP) * dim * dim);
        cec2013_synth_md(dim, m.data());
        d::MD.emplace(dim, std::move(m));
    }
    if (d::shift_data.empty()) {
        d::shift_data.resize(static_cast<std::size_t>(CEC_SYNTH_NCOMP) * 100);
        cec2013_synth_shift(d::shift_data.data());
    }
}

const std::vector<double> &cec2014_rotation(unsigned func, unsigned dim) { return pagmo::detail::cec2014_data::rotation_data.at(func).at(dim); }
const std::vector<double> &cec2014_shift(unsigned func) { return pagmo::detail::cec2014_data::shift_data.at(func); }
const std::vector<int> &cec2014_shuffle(unsigned func, unsigned dim) { return pagmo::detail::cec2014_data::shuffle_data.at(func).at(dim); }
const std::vector<double> &cec2013_md(unsigned dim) { return pagmo::detail::cec2013_data::MD.at(dim); }
const std::vector<double> &cec2013_shift() { return pagmo::detail::cec2013_data::shift_data; }

} // namespace oracle_ref
