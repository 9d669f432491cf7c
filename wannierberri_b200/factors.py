"""Physical prefactors multiplying the final arrays on the host (reference: factors.py:1-25),
from scipy.constants like the reference."""
from scipy.constants import elementary_charge, hbar, electron_mass, physical_constants, angstrom

bohr_magneton = elementary_charge * hbar / (2 * electron_mass)
bohr = physical_constants['Bohr radius'][0] / angstrom
eV_au = physical_constants['electron volt-hartree relationship'][0]
factor_morb_evA2_to_muB = -(elementary_charge * angstrom ** 2) * elementary_charge / (2 * hbar) / bohr_magneton
factor_gme_orb = factor_morb_evA2_to_muB * bohr_magneton / angstrom ** 2
factor_gme_spin = -bohr_magneton / angstrom ** 2
factor_ahc = -(elementary_charge ** 2 / hbar / angstrom)
factor_opt = -factor_ahc
factor_shc = -factor_ahc
TAU_UNIT = 1E-15
factor_ohmic = (elementary_charge ** 2 / hbar / angstrom * TAU_UNIT * elementary_charge / hbar)
factor_nlahc = elementary_charge ** 3 / hbar ** 2 * TAU_UNIT
fac_spin_Z = hbar / (2 * electron_mass)
factor_hall_classic = -(elementary_charge ** 3 / hbar ** 2 * angstrom * TAU_UNIT ** 2 * elementary_charge ** 2 / hbar ** 2)
factor_nldrude = -(elementary_charge ** 3 / hbar ** 2 * TAU_UNIT ** 2 * elementary_charge / hbar)
from math import pi  # noqa: E402
factor_shift_current = hbar / elementary_charge * pi * elementary_charge ** 3 / (4 * hbar ** 2)
factor_injection_current = -pi * elementary_charge ** 3 / (hbar ** 2) * TAU_UNIT
fac_orb_Z = elementary_charge / 2 / hbar * angstrom ** 2
factor_emcha = -(elementary_charge ** 4 / hbar ** 3 * angstrom ** 2 * TAU_UNIT ** 2 * elementary_charge / hbar)
