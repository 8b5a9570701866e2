"""ParameterReader contract of the drop-in API (reference src/ParameterReader.cpp:58-175):
`name = value  # comment` lines, names trimmed and lower-cased, values are doubles, later
assignments win, key=value command-line overrides, the shipped parameter files parse."""
import ctypes as C
import os

import cases
from iss_b200 import capi


def make(tmp_path, text):
    p = tmp_path/"params.dat"
    p.write_text(text)
    L = capi.host_lib()
    s = C.c_void_p(L.iss_host_create(str(tmp_path).encode(), capi.TABLES.encode(),
                                     capi.TABLES.encode(), str(p).encode(), b"surface.dat"))
    return L, s


def test_parsing_rules(built, tmp_path):
    L, s = make(tmp_path, "  Hydro_Mode = 2   # comment = 7\n"
                          "# whole line comment\n"
                          "\n"
                          "y_LB=-2.5\n"
                          "randomSeed   =   -1 #\n"
                          "hydro_mode = 1\n")
    g = lambda k, d=-99.0: L.iss_host_get_param(s, k.encode(), d)
    assert g("hydro_mode") == 1.0               # later assignment wins
    assert g("HYDRO_MODE") == 1.0               # names are case-insensitive
    assert g("y_lb") == -2.5
    assert g("randomseed") == -1.0
    assert g("comment") == -99.0                # text after '#' is ignored
    assert g("missing", 4.0) == 4.0
    L.iss_host_parse_param(s, b"Perform_Decays = 1 # from the command line")
    assert g("perform_decays") == 1.0
    L.iss_host_set_param(s, b"MC_sampling", 4.0)
    assert g("mc_sampling") == 4.0
    L.iss_host_destroy(s)


def test_shipped_parameter_files_parse(built, tmp_path):
    L = capi.host_lib()
    for name, expect in (("iSS_parameters.dat", 20.0), ("iSS_parameters_ideal.dat", 20.0),
                         ("iSS_parameters_CEdeltaf.dat", 21.0)):
        p = os.path.join(cases.FIX, name)
        s = C.c_void_p(L.iss_host_create(str(tmp_path).encode(), capi.TABLES.encode(),
                                         capi.TABLES.encode(), p.encode(), b"surface.dat"))
        g = lambda k: L.iss_host_get_param(s, k.encode(), -99.0)
        assert g("bulk_deltaf_kind") == expect
        assert g("mc_sampling") == 4.0
        assert g("number_of_repeated_sampling") == 2000.0
        for key in ("hydro_mode", "afterburner_type", "turn_on_bulk", "include_deltaf_shear",
                    "quantum_statistics", "dn_dy_sampling_model", "y_lb", "y_rb", "randomseed",
                    "maximum_sampling_events", "local_charge_conservation"):
            assert g(key) != -99.0, key
        L.iss_host_destroy(s)
