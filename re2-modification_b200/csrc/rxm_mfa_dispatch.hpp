// Compile-time instantiation table for the K2 MFA simulation:
//   NC   = cell slots carried per configuration (>= rxm_tables::n_cells)
//   CAP  = configuration slots per frontier      (>= n_states: one slot per node)
//   DMAX = depth of the explicit evaluateState recursion stack
// RXM_MFA_DISPATCH(nc, cap, CALL) expands CALL(NC, CAP, DMAX) for the smallest
// instantiation that fits, or sets `rxm_dispatch_ok = false`.
#ifndef RXM_MFA_DISPATCH_HPP
#define RXM_MFA_DISPATCH_HPP

#define RXM_MFA_DMAX 8
#define RXM_MFA_MAX_STATES 128

#define RXM_MFA_DISPATCH_CAP(NCV, cap, CALL)                 \
    do {                                                     \
        if ((cap) <= 8) { CALL(NCV, 8, RXM_MFA_DMAX); }      \
        else if ((cap) <= 16) { CALL(NCV, 16, RXM_MFA_DMAX); } \
        else if ((cap) <= 32) { CALL(NCV, 32, RXM_MFA_DMAX); } \
        else if ((cap) <= 64) { CALL(NCV, 64, RXM_MFA_DMAX); } \
        else if ((cap) <= 128) { CALL(NCV, 128, RXM_MFA_DMAX); } \
        else rxm_dispatch_ok = false;                        \
    } while (0)

#define RXM_MFA_DISPATCH(nc, cap, CALL)                         \
    do {                                                        \
        if ((nc) <= 2) RXM_MFA_DISPATCH_CAP(2, cap, CALL);      \
        else if ((nc) <= 4) RXM_MFA_DISPATCH_CAP(4, cap, CALL); \
        else if ((nc) <= 9) RXM_MFA_DISPATCH_CAP(9, cap, CALL); \
        else rxm_dispatch_ok = false;                           \
    } while (0)

#endif
