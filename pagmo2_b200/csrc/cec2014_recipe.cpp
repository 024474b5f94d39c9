// cec2014_recipe.cpp - host-side description ("recipe") of CEC2014 f1..f30 as stages of
//   shift -> scale -> (rotate) -> (permute) -> groups of primitives -> (scale) -> (cf_cal combine).
// The recipe is data, the device code in eval_cec2014.cu interprets it.  Every constant table is computed
// here with the HOST libm, i.e. the same libm the reference uses, so pow(10, 6i/(n-1)), sqrt(1+i),
// 2*pi*3^j ... are bit-identical to what reference src/problems/cec2014.cpp computes per call.
//
// Reference map (src/problems/cec2014.cpp): dispatch :119-247, primitives :375-783, hybrids hf01-06
// :786-1034, compositions cf01-08 :1037-1213, sr_func :1238-1274, cf_cal :1319-1353.
#include <cmath>
#include <cstring>

#include "pgc_internal.cuh"

namespace pgc
{

namespace
{

constexpr double kPi = 3.141592653589793238462643383279502884;

// sh_rate each primitive passes to sr_func (cec2014.cpp:381,393,407,437,474,498,522,539,573,602,749,772,700,724)
double prim_rate(int p)
{
    switch (p) {
        case P_ROSENBROCK: return 2.048 / 100.0;
        case P_WEIERSTRASS: return 0.5 / 100.0;
        case P_GRIEWANK: return 600.0 / 100.0;
        case P_RASTRIGIN: return 5.12 / 100.0;
        case P_SCHWEFEL: return 1000.0 / 100.0;
        case P_KATSUURA: return 5.0 / 100.0;
        case P_HAPPYCAT: return 5.0 / 100.0;
        case P_HGBAT: return 5.0 / 100.0;
        case P_GRIE_ROSEN: return 5.0 / 100.0;
        default: return 1.0; // ellips, bent_cigar, discus, ackley, escaffer6
    }
}

// FP64 add/mul per coordinate and libm calls per coordinate of each primitive (SURVEY.md 8d bookkeeping).
void prim_work(int p, double n, double &flops, double &transc)
{
    switch (p) {
        case P_ELLIPS: case P_BENT_CIGAR: case P_DISCUS: flops += 3 * n; break;
        case P_ROSENBROCK: flops += 8 * n; break;
        case P_ACKLEY: flops += 4 * n; transc += n + 3; break;
        case P_WEIERSTRASS: flops += 21 * 4 * n; transc += 21 * n; break;
        case P_GRIEWANK: flops += 4 * n; transc += n; break;
        case P_RASTRIGIN: flops += 5 * n; transc += n; break;
        case P_SCHWEFEL: flops += 8 * n; transc += 2 * n; break;
        case P_KATSUURA: flops += 32 * 5 * n; transc += n; break;
        case P_HAPPYCAT: case P_HGBAT: flops += 3 * n; transc += 2; break;
        case P_GRIE_ROSEN: flops += 14 * n; transc += n; break;
        case P_ESCAFFER6: flops += 10 * n; transc += 2 * n; break;
        default: break;
    }
}

void add_group(Cec2014Recipe &r, StageDesc &st, int prim, int off, int len, double rate)
{
    GroupDesc &g = st.g[st.ngroups++];
    g.prim = prim;
    g.off = off;
    g.len = len;
    g.rate = rate;
    g.tab_off = static_cast<int>(r.table.size());
    g.c0 = g.c1 = 0.0;
    const unsigned nx = static_cast<unsigned>(len);
    switch (prim) {
        case P_ELLIPS: // :383  pow(10.0, 6.0 * i / (nx - 1))
            for (unsigned i = 0; i < nx; ++i) r.table.push_back(std::pow(10.0, 6.0 * i / (nx - 1)));
            break;
        case P_GRIEWANK: // :526  z / sqrt(1.0 + i), kept as a reciprocal (<= 1 ulp from the true quotient)
            for (unsigned i = 0; i < nx; ++i) r.table.push_back(1.0 / std::sqrt(1.0 + i));
            break;
        case P_WEIERSTRASS: { // :491-509
            const double a = 0.5, b = 3.0;
            double sum2 = 0.0;
            for (unsigned j = 0; j <= 20; ++j) {
                r.table.push_back(2.0 * kPi * std::pow(b, j)); // argument scale of term j
                sum2 += std::pow(a, j) * std::cos(2.0 * kPi * std::pow(b, j) * 0.5);
            }
            for (unsigned j = 0; j <= 20; ++j) r.table.push_back(std::pow(a, j));
            g.c0 = nx * sum2; // :509  f -= nx * sum2
            break;
        }
        case P_SCHWEFEL: g.c0 = 4.189828872724338e+002 * nx; break; // :589
        case P_KATSUURA: {                                           // :600,611,613
            const double tmp3 = std::pow(1.0 * nx, 1.2);
            g.c0 = 10.0 / tmp3;
            g.c1 = 10.0 / nx / nx;
            break;
        }
        default: break;
    }
    // keep every table slice 16-byte aligned
    if (r.table.size() & 1u) r.table.push_back(0.0);
    prim_work(prim, len, r.flops_per_eval, r.transc_per_eval);
}

StageDesc &new_stage(Cec2014Recipe &r, int comp, int rotate, double pre_rate)
{
    StageDesc &st = r.st[r.nstages++];
    std::memset(&st, 0, sizeof(st));
    st.comp = comp;
    st.rotate = rotate;
    st.pre_rate = pre_rate;
    const double D = r.dim;
    r.flops_per_eval += 2 * D + (rotate ? 2 * D * D : 0);
    return st;
}

// basic function as a stage: sr_func(x, z, nx, Os, Mr, rate, 1, r_flag) then the primitive on z
void basic_stage(Cec2014Recipe &r, int comp, int prim, int rotate)
{
    StageDesc &st = new_stage(r, comp, rotate, prim_rate(prim));
    add_group(r, st, prim, 0, r.dim, 1.0);
}

// hybrid hfXX as a stage (:786-1034): sr_func(rate 1.0, rotate) -> y[j] = z[S[j]-1] -> groups
void hybrid_stage(Cec2014Recipe &r, int comp, int rotate, int cf_num, const double *Gp, const int *prims)
{
    StageDesc &st = new_stage(r, comp, rotate, 1.0);
    st.permute = 1;
    const unsigned nx = static_cast<unsigned>(r.dim);
    unsigned G_nx[kMaxGroups], tmp = 0;
    for (int i = 0; i < cf_num - 1; ++i) { // :795-799
        G_nx[i] = static_cast<unsigned>(std::ceil(Gp[i] * nx));
        tmp += G_nx[i];
    }
    G_nx[cf_num - 1] = nx - tmp;
    unsigned off = 0;
    for (int i = 0; i < cf_num; ++i) {
        add_group(r, st, prims[i], static_cast<int>(off), static_cast<int>(G_nx[i]), prim_rate(prims[i]));
        off += G_nx[i];
    }
}

void hybrid_by_id(Cec2014Recipe &r, int comp, int rotate, int hf)
{
    static const double Gp3[3] = {0.3, 0.3, 0.4};
    static const double Gp4[4] = {0.2, 0.2, 0.3, 0.3};
    static const double Gp5[5] = {0.1, 0.2, 0.2, 0.2, 0.3};
    static const int p1[3] = {P_SCHWEFEL, P_RASTRIGIN, P_ELLIPS};                              // hf01 :811-815
    static const int p2[3] = {P_BENT_CIGAR, P_HGBAT, P_RASTRIGIN};                             // hf02 :849-853
    static const int p3[4] = {P_GRIEWANK, P_WEIERSTRASS, P_ROSENBROCK, P_ESCAFFER6};           // hf03 :889-895
    static const int p4[4] = {P_HGBAT, P_DISCUS, P_GRIE_ROSEN, P_RASTRIGIN};                   // hf04 :931-937
    static const int p5[5] = {P_ESCAFFER6, P_HGBAT, P_ROSENBROCK, P_SCHWEFEL, P_ELLIPS};       // hf05 :975-983
    static const int p6[5] = {P_KATSUURA, P_HAPPYCAT, P_GRIE_ROSEN, P_SCHWEFEL, P_ACKLEY};     // hf06 :1021-1029
    switch (hf) {
        case 1: hybrid_stage(r, comp, rotate, 3, Gp3, p1); break;
        case 2: hybrid_stage(r, comp, rotate, 3, Gp3, p2); break;
        case 3: hybrid_stage(r, comp, rotate, 4, Gp4, p3); break;
        case 4: hybrid_stage(r, comp, rotate, 4, Gp4, p4); break;
        case 5: hybrid_stage(r, comp, rotate, 5, Gp5, p5); break;
        case 6: hybrid_stage(r, comp, rotate, 5, Gp5, p6); break;
    }
}

void scale_last(Cec2014Recipe &r, double mul, double div)
{
    StageDesc &st = r.st[r.nstages - 1];
    st.scaled = 1;
    st.mul = mul;
    st.div = div;
    r.flops_per_eval += 2;
}

void composition(Cec2014Recipe &r, int n, const double *delta, const double *bias)
{
    r.composition = 1;
    for (int i = 0; i < n; ++i) {
        r.delta[i] = delta[i];
        r.cbias[i] = bias[i];
    }
    // cf_cal :1327-1351: per component nx*(sub,mul,add) + ~8, one exp, one pow(.,0.5)
    r.flops_per_eval += n * (3.0 * r.dim + 8);
    r.transc_per_eval += 2.0 * n;
}

} // namespace

int build_cec2014_recipe(unsigned func, unsigned dim, Cec2014Recipe &r)
{
    if (!(dim == 2u || dim == 10u || dim == 20u || dim == 30u || dim == 50u || dim == 100u)) { // :51-55
        set_error("Error: CEC2014 Test functions are only defined for dimensions 2,10,20,30,50,100, a dimension of %u "
                  "was detected.",
                  dim);
        return PGC_ERR_INVALID_ARGUMENT;
    }
    if (func < 1u || func > 30u) { // :56-60
        set_error("Error: CEC2014 Test functions are only defined for prob_id in [1, 28], a prob_id of %u was detected.",
                  func);
        return PGC_ERR_INVALID_ARGUMENT;
    }
    if (dim == 2u && ((func >= 17u && func <= 22u) || (func >= 29u && func <= 30u))) { // :62-64
        set_error("hf01,hf02,hf03,hf04,hf05,hf06,cf07&cf08 are NOT defined for D=2.");
        return PGC_ERR_INVALID_ARGUMENT;
    }
    r.func = static_cast<int>(func);
    r.dim = static_cast<int>(dim);
    r.nstages = 0;
    r.composition = 0;
    r.fbias = 100.0 * func;
    r.table.clear();
    r.flops_per_eval = 1; // + bias
    r.transc_per_eval = 0;

    static const double d1[5] = {10, 20, 30, 40, 50}, b5[5] = {0, 100, 200, 300, 400};
    static const double d20[3] = {20, 20, 20}, d135[3] = {10, 30, 50}, b3[3] = {0, 100, 200};
    static const double d10[5] = {10, 10, 10, 10, 10}, d5[5] = {10, 10, 10, 20, 20};

    switch (func) {
        case 1: basic_stage(r, 0, P_ELLIPS, 1); break;
        case 2: basic_stage(r, 0, P_BENT_CIGAR, 1); break;
        case 3: basic_stage(r, 0, P_DISCUS, 1); break;
        case 4: basic_stage(r, 0, P_ROSENBROCK, 1); break;
        case 5: basic_stage(r, 0, P_ACKLEY, 1); break;
        case 6: basic_stage(r, 0, P_WEIERSTRASS, 1); break;
        case 7: basic_stage(r, 0, P_GRIEWANK, 1); break;
        case 8: basic_stage(r, 0, P_RASTRIGIN, 0); break;
        case 9: basic_stage(r, 0, P_RASTRIGIN, 1); break;
        case 10: basic_stage(r, 0, P_SCHWEFEL, 0); break;
        case 11: basic_stage(r, 0, P_SCHWEFEL, 1); break;
        case 12: basic_stage(r, 0, P_KATSUURA, 1); break;
        case 13: basic_stage(r, 0, P_HAPPYCAT, 1); break;
        case 14: basic_stage(r, 0, P_HGBAT, 1); break;
        case 15: basic_stage(r, 0, P_GRIE_ROSEN, 1); break;
        case 16: basic_stage(r, 0, P_ESCAFFER6, 1); break;
        case 17: case 18: case 19: case 20: case 21: case 22: hybrid_by_id(r, 0, 1, static_cast<int>(func) - 16); break;
        case 23: // cf01 :1037-1061
            basic_stage(r, 0, P_ROSENBROCK, 1); scale_last(r, 10000, 1e+4);
            basic_stage(r, 1, P_ELLIPS, 1); scale_last(r, 10000, 1e+10);
            basic_stage(r, 2, P_BENT_CIGAR, 1); scale_last(r, 10000, 1e+30);
            basic_stage(r, 3, P_DISCUS, 1); scale_last(r, 10000, 1e+10);
            basic_stage(r, 4, P_ELLIPS, 0); scale_last(r, 10000, 1e+10);
            composition(r, 5, d1, b5);
            break;
        case 24: // cf02 :1064-1079
            basic_stage(r, 0, P_SCHWEFEL, 0);
            basic_stage(r, 1, P_RASTRIGIN, 1);
            basic_stage(r, 2, P_HGBAT, 1);
            composition(r, 3, d20, b3);
            break;
        case 25: // cf03 :1082-1099
            basic_stage(r, 0, P_SCHWEFEL, 1); scale_last(r, 1000, 4e+3);
            basic_stage(r, 1, P_RASTRIGIN, 1); scale_last(r, 1000, 1e+3);
            basic_stage(r, 2, P_ELLIPS, 1); scale_last(r, 1000, 1e+10);
            composition(r, 3, d135, b3);
            break;
        case 26: // cf04 :1102-1125
            basic_stage(r, 0, P_SCHWEFEL, 1); scale_last(r, 1000, 4e+3);
            basic_stage(r, 1, P_HAPPYCAT, 1); scale_last(r, 1000, 1e+3);
            basic_stage(r, 2, P_ELLIPS, 1); scale_last(r, 1000, 1e+10);
            basic_stage(r, 3, P_WEIERSTRASS, 1); scale_last(r, 1000, 400);
            basic_stage(r, 4, P_GRIEWANK, 1); scale_last(r, 1000, 100);
            composition(r, 5, d10, b5);
            break;
        case 27: // cf05 :1128-1151
            basic_stage(r, 0, P_HGBAT, 1); scale_last(r, 10000, 1000);
            basic_stage(r, 1, P_RASTRIGIN, 1); scale_last(r, 10000, 1e+3);
            basic_stage(r, 2, P_SCHWEFEL, 1); scale_last(r, 10000, 4e+3);
            basic_stage(r, 3, P_WEIERSTRASS, 1); scale_last(r, 10000, 400);
            basic_stage(r, 4, P_ELLIPS, 1); scale_last(r, 10000, 1e+10);
            composition(r, 5, d5, b5);
            break;
        case 28: // cf06 :1154-1177
            basic_stage(r, 0, P_GRIE_ROSEN, 1); scale_last(r, 10000, 4e+3);
            basic_stage(r, 1, P_HAPPYCAT, 1); scale_last(r, 10000, 1e+3);
            basic_stage(r, 2, P_SCHWEFEL, 1); scale_last(r, 10000, 4e+3);
            basic_stage(r, 3, P_ESCAFFER6, 1); scale_last(r, 10000, 2e+7);
            basic_stage(r, 4, P_ELLIPS, 1); scale_last(r, 10000, 1e+10);
            composition(r, 5, d1, b5);
            break;
        case 29: // cf07 :1180-1195
            hybrid_by_id(r, 0, 1, 1);
            hybrid_by_id(r, 1, 1, 2);
            hybrid_by_id(r, 2, 1, 3);
            composition(r, 3, d135, b3);
            break;
        case 30: // cf08 :1198-1213
            hybrid_by_id(r, 0, 1, 4);
            hybrid_by_id(r, 1, 1, 5);
            hybrid_by_id(r, 2, 1, 6);
            composition(r, 3, d135, b3);
            break;
    }
    return PGC_OK;
}

} // namespace pgc
