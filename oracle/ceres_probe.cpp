// TEST INFRASTRUCTURE ONLY -- driver around the UNMODIFIED reference source
//   lib/csrc/uncertainty_pnp/src/uncertainty_pnp.cpp   (functor :7-55, C entry `uncertainty_pnp` :61-92)
// linked against the reference's own prebuilt lib/libceres.so.2.0.0 (oracle/build_ceres_ref.py).
//
// The reference file is #included where it lies under /root/reference (never copied into this repository): the
// translation unit therefore contains the reference's `uncertainty_pnp(...)` itself -- the function the product's
// pvb_uncertainty_pnp replaces -- plus `ceres_probe(...)`, which builds the SAME ceres::Problem with the SAME options
// (uncertainty_pnp.cpp:72-88) and additionally reports what the C entry throws away: Solver::Summary and the
// per-iteration trace.  Used to pin oracle/pnp_oracle.py and the CUDA kernel against real Ceres iterates.
#define main pvb_reference_demo_main          // uncertainty_pnp.cpp:98 carries a demo main()
#include "uncertainty_pnp.cpp"
#undef main
#include <cstring>
#include <vector>

namespace {
struct TraceCallback : public ceres::IterationCallback {
    std::vector<double> *rows;
    explicit TraceCallback(std::vector<double> *r) : rows(r) {}
    ceres::CallbackReturnType operator()(const ceres::IterationSummary &s) override
    {
        const double row[8] = {(double)s.iteration, s.cost, s.cost_change, s.gradient_max_norm, s.step_norm,
                               s.relative_decrease, s.trust_region_radius, s.step_is_successful ? 1.0 : 0.0};
        rows->insert(rows->end(), row, row + 8);
        return ceres::SOLVER_CONTINUE;
    }
};
}

static char g_last_message[512] = "";

extern "C" {

const char *ceres_probe_last_message(void) { return g_last_message; }

// summary8 = {termination_type, #IterationSummary, initial_cost, final_cost, successful, unsuccessful steps,
//             linear_solver_type_used, stop reason (codes below)};  trace = [rows][8] as in TraceCallback; returns the number of rows written.
int ceres_probe(double *pts2d, double *pts3d, double *wgt2d, double *K, double *init_rt, double *result_rt, int pn,
                double *summary8, double *trace, int max_rows)
{
    ceres::Problem problem;
    double solution[6];
    memcpy(solution, init_rt, sizeof(solution));
    for (int i = 0; i < pn; ++i)
        problem.AddResidualBlock(
            ReprojectionErrorArray::Create(pts2d[i * 2], pts2d[i * 2 + 1], pts3d[i * 3], pts3d[i * 3 + 1], pts3d[i * 3 + 2],
                                           wgt2d[i * 3], wgt2d[i * 3 + 1], wgt2d[i * 3 + 2], K[0], K[4], K[2], K[5]),
            NULL, solution);
    ceres::Solver::Options options;
    options.linear_solver_type = ceres::DENSE_SCHUR;
    std::vector<double> rows;
    TraceCallback cb(&rows);
    options.callbacks.push_back(&cb);
    ceres::Solver::Summary summary;
    ceres::Solve(options, &problem, &summary);
    memcpy(result_rt, solution, sizeof(solution));
    summary8[0] = (double)summary.termination_type;
    summary8[1] = (double)summary.iterations.size();
    summary8[2] = summary.initial_cost;
    summary8[3] = summary.final_cost;
    summary8[4] = (double)summary.num_successful_steps;
    summary8[5] = (double)summary.num_unsuccessful_steps;
    summary8[6] = (double)summary.linear_solver_type_used;
    // why it stopped, from Summary::message (trust_region_minimizer.cc): 1 gradient, 2 parameter, 3 function tolerance,
    // 4 trust-region radius, 5 iteration cap, 6 anything else (failure / invalid steps), same codes as pvb_uncertainty_pnp's info
    const std::string &m = summary.message;
    double why = 6.0;
    if (m.find("Gradient tolerance reached") != std::string::npos) why = 1.0;
    else if (m.find("Parameter tolerance reached") != std::string::npos) why = 2.0;
    else if (m.find("Function tolerance reached") != std::string::npos) why = 3.0;
    else if (m.find("trust region radius") != std::string::npos || m.find("Trust region radius") != std::string::npos) why = 4.0;
    else if (m.find("Maximum number of iterations reached") != std::string::npos) why = 5.0;
    summary8[7] = why;
    strncpy(g_last_message, m.c_str(), sizeof(g_last_message) - 1);
    int n = (int)(rows.size() / 8);
    if (n > max_rows) n = max_rows;
    if (n > 0) memcpy(trace, rows.data(), sizeof(double) * 8 * (size_t)n);
    return n;
}

}   // extern "C"
