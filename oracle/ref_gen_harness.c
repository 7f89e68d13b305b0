/* oracle/ref_gen_harness.c -- ft8_lib's signal generator (gen_ft8.c) as a shared object.  TEST INFRASTRUCTURE ONLY.
 * The reference translation unit is textually included from where it lies under /root/reference (-I$(REF)/ft8_lib), main()
 * renamed; nothing of it is copied into this repository.  Used to compare the GFSK mode of the synthesiser (csrc/synth.cu and its
 * CPU twin ft8_oracle_synth.c) with the reference's own gfsk_pulse() / synth_gfsk() (gen_ft8.c:28-102). */
#define main ref_gen_main
#include "gen_ft8.c"
#undef main

void refgen_pulse(int n_spsym, float symbol_bt, float *pulse) { gfsk_pulse(n_spsym, symbol_bt, pulse); }
void refgen_synth(const uint8_t *symbols, int n_sym, float f0, float symbol_bt, float symbol_period, int signal_rate, float *signal) {
    synth_gfsk(symbols, n_sym, f0, symbol_bt, symbol_period, signal_rate, signal);
}
