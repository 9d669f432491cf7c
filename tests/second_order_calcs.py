"""The second-order calculators of the parity fixture, created from the reference's `calculators.static` module
(tests/golden/make_golden_second_order.py, which makes the fixture) or from this package's (tests/test_gpu_parity.py,
tests/test_host.py): the same calls on both."""


def make_calculators(st, Efermi):
    """the same calls on the reference's module and on this package's (tests/test_gpu_parity.py)"""
    ff = dict(FF_rotAA=True)
    return dict(
        z_spin=st.NLDrude_Zeeman_spin(Efermi=Efermi),
        z_orb_omega=st.NLDrude_Zeeman_orb_Omega(Efermi=Efermi),
        z_orb=st.NLDrude_Zeeman_orb(Efermi=Efermi),
        emcha=st.eMChA_FermiSurf(Efermi=Efermi),
        qmetric=st.QuantumMetric_FermiSea(Efermi=Efermi, kwargs_formula=ff),
        qmetric_dip=st.QuantumMetric_Vel_DQ(Efermi=Efermi, kwargs_formula=ff),
        z_orb_int=st.NLDrude_Zeeman_orb(Efermi=Efermi, kwargs_formula=dict(external_terms=False)),
        emcha_int=st.eMChA_FermiSurf(Efermi=Efermi, kwargs_formula=dict(external_terms=False)),
        qmetric_int=st.QuantumMetric_FermiSea(Efermi=Efermi, kwargs_formula=dict(external_terms=False)),
        z_spin_wide=st.NLDrude_Zeeman_spin(Efermi=Efermi, degen_thresh=0.3, degen_Kramers=True),
        emcha_wide=st.eMChA_FermiSurf(Efermi=Efermi, degen_thresh=0.3, degen_Kramers=True, use_factor=False),
        qmetric_wide=st.QuantumMetric_FermiSea(Efermi=Efermi, kwargs_formula=ff, degen_thresh=0.3),
    )
