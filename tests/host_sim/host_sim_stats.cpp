// host_sim_stats.cpp -- TEST INFRASTRUCTURE: storage for the optional event counters of csrc/a26_core.cuh (A26_STATS builds of
// host_sim.cpp only).  Index 5 = dispatcher trips, 6 / 7 / 8 = calls of the main display loop / score loop / blank-line loop super-blocks.
unsigned long long a26_stats[16];
unsigned long long a26_entry_stats[2048];
unsigned long long a26_reg_stats[64];
extern "C" unsigned long long *hs_stats() { return a26_stats; }
