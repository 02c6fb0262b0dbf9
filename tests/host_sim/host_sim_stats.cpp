// host_sim_stats.cpp -- TEST INFRASTRUCTURE: storage for the optional event counters of csrc/a26_core.cuh (A26_STATS builds of
// host_sim.cpp only).  Index 5 = dispatcher trips, 6 = main display loop super-block calls, 7 = score loop super-block calls.
unsigned long long a26_stats[16];
unsigned long long a26_entry_stats[2048];
unsigned long long a26_reg_stats[64];
extern "C" unsigned long long *hs_stats() { return a26_stats; }
