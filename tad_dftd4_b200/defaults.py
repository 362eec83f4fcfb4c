"""Default values (same numbers as the reference's ``defaults.py:26-94``)."""

D4_CN_CUTOFF = 30.0
D4_CN_EEQ_CUTOFF = 25.0
D4_CN_EEQ_MAX = 8.0
D4_DISP2_CUTOFF = 60.0
D4_DISP3_CUTOFF = 40.0
D4_KCN = 7.5
D4_K4 = 4.10451
D4_K5 = 19.08857
D4_K6 = 2 * 11.28174**2
A1 = 0.4
A2 = 5.0
RS6 = 1.0
S6 = 1.0
RS8 = 1.0
S8 = 1.0
S9 = 1.0
S10 = 0.0
RS9 = 1.0
ALP = 16.0
BET = 0.0
GA_DEFAULT = 3.0
GC_DEFAULT = 2.0
WF_DEFAULT = 6.0
