// results_host.h -- host pieces of results() shared by cd_results_adjust (results.cpp) and the device-resident
// variant orchestrated in context.cu
#pragma once

namespace cd {

double res_qf(double prob, double df1, double df2);                          // stats::qf by bisection on pbeta
int res_pick_cutoff(const double* theta, const double* numRej, int nt);     // lowess + threshold rule -> index

}  // namespace cd
