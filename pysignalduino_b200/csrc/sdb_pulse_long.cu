/*
 * sdb_pulse_long.cu — the MS / MU kernels of sdb_pulse.cu compiled a second time, sized for messages of
 * SDB_FAST_DIGITS < D <= SDB_MAX_DIGITS digits (namespace sdb_long; 2 warps per CTA; resolve + fused scan).
 * The reference puts no limit on len(D) (sd_protocols/message_unsynced.py:22-25, signalduino/parser/mu.py:48).
 */
#define SDB_PULSE_LONG 1
#include "sdb_pulse.cu"
