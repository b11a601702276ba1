// TEST INFRASTRUCTURE ONLY (see oracle/oracle.h): CPU restatement of the reference's grain diagnostics,
// Grains::findAndNumberGrains (source/Grains.cc:263-520) and Grains::computeGrainVolumes (Grains.cc:647-697),
// on one uniform level: the reference's own algorithm -- unique cell numbers, then in-place sweeps in cell order that
// take the lowest number among the face neighbours, ghost numbers refilled before every sweep (periodic directions
// wrap, physical boundaries keep -1: do_physical_boundary_fill = false), until a sweep changes nothing.
// Self-contained (plain arrays in, no Ctx), so the long-double build does not need it.
#include <map>
#include <vector>

extern "C" int oracle_grain_volumes(int ndim, const int* n_, const double* dx, const int* zero_slope, const double* phase,
                                    double phase_threshold, int max_grains, int* ngrains, int* grain_ids,
                                    double* volumes, int* grain_number_out)
{
   int n[3] = {n_[0], n_[1], ndim == 3 ? n_[2] : 1};
   const int g[3] = {n[0] + 2, n[1] + 2, ndim == 3 ? n[2] + 2 : 1};  // ghost width 1 (no ghost along a missing axis)
   const int oz = (ndim == 3) ? 1 : 0;
   auto at = [&](int i, int j, int k) { return (size_t)(i + 1) + (size_t)g[0] * ((size_t)(j + 1) + (size_t)g[1] * (size_t)(k + oz)); };
   std::vector<int> num((size_t)g[0] * g[1] * g[2], -1);  // g->fillAll(-1, ghost box), Grains.cc:346
   // unique number per cell: level_stride (Grains.cc:307-312, 353-358)
   for (int k = 0; k < n[2]; k++)
      for (int j = 0; j < n[1]; j++)
         for (int i = 0; i < n[0]; i++)
            if (phase[(size_t)i + (size_t)n[0] * ((size_t)j + (size_t)n[1] * k)] >= phase_threshold)
               num[at(i, j, k)] = i + n[0] * (j + n[1] * k);
   int width = 0;
   for (int d = 0; d < ndim; d++) width = (n[d] > width) ? n[d] : width;
   const int max_iteration_count = 2 * 2 * width;  // Grains.cc:314-324
   auto fill_ghosts = [&]() {  // refine schedule fillData without physical boundary fill (Grains.cc:391-394)
      for (int d = 0; d < ndim; d++) {
         if (zero_slope[d]) continue;  // physical boundary: ghosts stay -1
         for (int k = (d == 2 ? 0 : -oz); k < n[2] + (d == 2 ? 0 : oz); k++)
            for (int j = (d == 1 ? 0 : -1); j < n[1] + (d == 1 ? 0 : 1); j++)
               for (int i = (d == 0 ? 0 : -1); i < n[0] + (d == 0 ? 0 : 1); i++) {
                  // one layer below and above along d, copied from the opposite interior layer
                  if (d == 0 && i == 0) num[at(-1, j, k)] = num[at(n[0] - 1, j, k)], num[at(n[0], j, k)] = num[at(0, j, k)];
                  if (d == 1 && j == 0) num[at(i, -1, k)] = num[at(i, n[1] - 1, k)], num[at(i, n[1], k)] = num[at(i, 0, k)];
                  if (d == 2 && k == 0) num[at(i, j, -1)] = num[at(i, j, n[2] - 1)], num[at(i, j, n[2])] = num[at(i, j, 0)];
               }
      }
   };
   int counter = 0;
   while (counter < max_iteration_count) {  // Grains.cc:383-517
      int changed = 0;
      fill_ghosts();
      for (int k = 0; k < n[2]; k++)
         for (int j = 0; j < n[1]; j++)
            for (int i = 0; i < n[0]; i++) {
               const int nn = num[at(i, j, k)];
               if (nn < 0) continue;  // (the weight of every cell of a uniform level is > 0)
               for (int dd = 0; dd < ndim; dd++) {  // Grains.cc:483-503: compared with the number read before the loop
                  const int di = dd == 0, dj = dd == 1, dk = dd == 2;
                  const int nm = num[at(i - di, j - dj, k - dk)];
                  if (nm >= 0 && nm < nn) num[at(i, j, k)] = nm, changed = 1;
                  const int np = num[at(i + di, j + dj, k + dk)];
                  if (np >= 0 && np < nn) num[at(i, j, k)] = np, changed = 1;
               }
            }
      counter++;
      if (!changed) break;
   }
   // computeGrainVolumes (Grains.cc:647-697): std::map<int, double> += control volume, cell by cell
   double dv = 1.0;
   for (int d = 0; d < ndim; d++) dv *= dx[d];
   std::map<int, double> vol;
   for (int k = 0; k < n[2]; k++)
      for (int j = 0; j < n[1]; j++)
         for (int i = 0; i < n[0]; i++) {
            const int nn = num[at(i, j, k)];
            if (grain_number_out) grain_number_out[(size_t)i + (size_t)n[0] * ((size_t)j + (size_t)n[1] * k)] = nn;
            if (nn >= 0 && dv > 0.) vol[nn] += dv;
         }
   *ngrains = (int)vol.size();
   if ((int)vol.size() > max_grains) return -1;
   int m = 0;
   for (const auto& it : vol) grain_ids[m] = it.first, volumes[m] = it.second, m++;
   return 0;
}
