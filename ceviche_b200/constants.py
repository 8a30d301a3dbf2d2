"""Physical constants with the reference's exact digits (ceviche/constants.py:7-11).
They are NOT the SI values; C_0 is derived and equals 299792458.13099605.  dt, the PML
sigma and every update coefficient depend on them, so they are reproduced digit for digit."""
from math import sqrt

EPSILON_0 = 8.85418782e-12
MU_0 = 1.25663706e-6
C_0 = 1 / sqrt(EPSILON_0 * MU_0)
ETA_0 = sqrt(MU_0 / EPSILON_0)
Q_e = 1.602176634e-19
