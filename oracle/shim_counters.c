/* oracle/shim_counters.c -- TEST INFRASTRUCTURE: storage for the cblas shim's call counter
 * when the reference CLI (train.cpp) is linked without the harness. */
long ocffm_shim_dscal_calls = 0;
