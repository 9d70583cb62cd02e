// results_host.h -- host pieces of results() shared by cd_results_adjust (results.cpp) and the device-resident
// variant orchestrated in context.cu
#pragma once
#include <cstdint>
#include <vector>

namespace cd {

double res_qf(double prob, double df1, double df2);                          // stats::qf by bisection on pbeta
int res_pick_cutoff(const double* theta, const double* numRej, int nt);     // lowess + threshold rule -> index
// IHWcorrection's breaks <- (c(minLogDist, Inf) + c(0, maxLogDist)) / 2 (chicdiff.R:2039), sorted as cut() sorts them;
// false when one is NaN or two coincide ('breaks' are not unique)
bool ihw_breaks(int ngroups, const double* minLogDist, const double* maxLogDist, std::vector<double>& breaks);
// mean(out$avWeights) over the n rows from the per-group row counts (per_group[0] = rows without a group, which poison
// it), as R's two-pass long-double mean over the rows ordered by stratum computes it
double ihw_mean_weight(int64_t n, int ngroups, const unsigned long long* per_group, const double* avWeights);

}  // namespace cd
