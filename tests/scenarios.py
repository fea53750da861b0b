"""Fit scenarios shared by the golden-vector generator and the parity tests.

C1 scenarios restate the sections of the reference's own LM-solver test file
(reference c++/tests/lm_solver.cpp); each carries the known-answer values that
file asserts (chi2, tau, I0-0, bgr-0, I0-1, bgr-1) with the tolerance it uses.
C2-C5 are the BASELINE configs scaled down to sizes the dense reference can
run.
"""
from __future__ import annotations

import numpy as np

from gadfit_b200 import workloads as W

# index into fix_d (c++/tests/fixtures.h:9-17)
I0_0, BGR_0, TAU, I0_1, BGR_1 = 0, 1, 3, 4, 5


def _c1(name, lines, active, kat, settings=None, tol=1e-14, i0=(I0_0, I0_1), tau_name=None,
        n_points=(100, 100)):
    """active = (I0-0, bgr-0, I0-1, bgr-1, tau) activity flags."""
    return dict(name=name, kind="c1", lines=lines, active=active, kat=kat, tol=tol,
                settings=dict(iteration_limit=4) | (settings or {}), lambda0=1.0, i0=i0,
                n_points=n_points)


# (chi2, tau, I0-0, bgr-0, I0-1, bgr-1); None = "stays at its fixed start value"
C1_SCENARIOS = [
    _c1("indexing_all", "21-47", (1, 1, 1, 1, 1),
        (11620.0867270475, 17.8650243622964, 39.77705004578393, 13.57729652858559,
         129.0275065609783, 16.09079665934463)),
    _c1("indexing_bgr0_bgr1", "48-64", (0, 1, 0, 1, 1),
        (153628.8903849508, 31.95892116514992, None, 17.81484199806565, None, 36.73244337347508)),
    _c1("indexing_I00_I01", "65-81", (1, 0, 1, 0, 1),
        (10810.65153981582, 21.30228862988602, 56.42893238415446, None, 139.4901380914605, None)),
    _c1("indexing_tau_only", "82-96", (0, 0, 0, 0, 1),
        (51624.83919460665, 10.99329301695744, None, None, None, None), i0=(16, 17)),
    _c1("indexing_bgr0_I01_bgr1", "97-114", (0, 1, 1, 1, 1),
        (15974.61260816282, 20.47926391663428, None, 18.47600900933105, 143.0431252627765,
         9.453915929181857)),
    _c1("indexing_I00_bgr0_bgr1", "115-132", (1, 1, 0, 1, 1),
        (145780.4588072044, 8.408237957600141, 45.87087327322397, 16.59126759913267, None,
         36.38255403506549)),
    _c1("indexing_I00_I01_bgr1", "133-150", (1, 0, 1, 1, 1),
        (11623.17388899667, 20.61333132315124, 56.5139576021328, None, 134.8973104943701,
         11.77612256514583)),
    _c1("indexing_I00_bgr0_I01", "151-168", (1, 1, 1, 0, 1),
        (30610.67204238365, 16.54682323514368, 29.98632400541692, 12.99477135618182,
         124.6991105597198, None)),
    _c1("indexing_I00_bgr1", "169-185", (1, 0, 0, 1, 1),
        (150672.9869101836, 16.73368044360274, 53.73848940201638, None, None, 36.50405720192947)),
    _c1("indexing_bgr0_I01", "186-202", (0, 1, 1, 0, 1),
        (15348.60122706107, 21.87456778662339, None, 18.39176693290169, 147.1783948678938, None)),
    _c1("no_iterations", "379-393", (1, 1, 1, 1, 1),
        (284681.4650859562, 0.5356792380861322, 6.13604207015635, 2.960644474827888,
         -1.472720596147903, 4.251266120087788), settings=dict(iteration_limit=0)),
    _c1("no_iteration_limit", "394-411", (1, 1, 1, 1, 1),
        (5640.175130917765, 20.85609539787557, 46.44788540145462, 10.32140443380387,
         152.2711588123377, 5.533936911147024), settings=dict(iteration_limit=100), tol=1e-8),
    _c1("dtd_no_constraints", "427-443", (1, 1, 1, 1, 1),
        (5761.320550200902, 20.15808767822605, 46.33952661438988, 10.40696426744941,
         151.8449704289686, 5.793008018375257),
        settings=dict(iteration_limit=5, damp_max=False, DTD_min=[2.0, 1.0, 3.0, 1.0, 7.0])),
    _c1("dtd_default", "444-457", (1, 1, 1, 1, 1),
        (5687.451130305415, 21.01892108898218, 46.18357253310398, 10.48386354002993,
         151.5283959798012, 6.087406702661871), settings=dict(iteration_limit=5)),
    _c1("dtd_min", "458-472", (1, 1, 1, 1, 1),
        (5640.44344354764, 20.8194135048056, 46.44560357121117, 10.33411834275264,
         152.2267670731584, 5.560463422860977),
        settings=dict(iteration_limit=5, DTD_min=[2.0, 1.0, 3.0, 1.0, 7.0])),
    _c1("geodesic_acceleration", "475-497", (1, 1, 1, 1, 1),
        (5641.66030550462, 20.70654799943915, 46.48065799723029, 10.39142422387268,
         152.4514268293043, 5.748941149916498),
        settings=dict(iteration_limit=5, acceleration_threshold=0.9)),
    _c1("loss_linear", "512-526", (1, 1, 1, 1, 1),
        (5687.451130305415, 21.01892108898218, 46.18357253310398, 10.48386354002993,
         151.5283959798012, 6.087406702661871), settings=dict(iteration_limit=5, loss=0)),
    _c1("loss_cauchy", "527-540", (1, 1, 1, 1, 1),
        (16869.67716299524, 17.45448014750576, 40.28201426242013, 9.278480584355261,
         132.6242198264016, 6.7051221338403), settings=dict(iteration_limit=5, loss=1)),
    _c1("loss_huber", "541-556", (1, 1, 1, 1, 1),
        (123695.8709974329, 4.643243104460152, 52.6348486049053, 7.874003370245958,
         166.3872296081963, 7.690335499679898), settings=dict(iteration_limit=2, loss=2)),
    _c1("no_degrees_of_freedom", "323-342", (1, 1, 1, 1, 1),
        (None, 2.945868346541778, 7.351966871429338, 49.68674387147222, -13.18731292934346,
         162.1781165060037), tol=1e-12, n_points=(3, 2)),
]

# BASELINE configs 2-5 at sizes the dense reference runs in well under a second
SCALED_SCENARIOS = [
    dict(name="C2_scaled", kind="workload", workload="C2", kwargs=dict(n_sets=4, n_points=2000)),
    dict(name="C3_scaled", kind="workload", workload="C3", kwargs=dict(n_sets=8, n_points=512)),
    dict(name="C3_scaled_40", kind="workload", workload="C3", kwargs=dict(n_sets=40, n_points=300)),
    dict(name="C4_scaled", kind="workload", workload="C4", kwargs=dict(n_sets=4, n_points=256)),
    dict(name="C5_scaled", kind="workload", workload="C5", kwargs=dict(n_sets=6, n_points=64)),
    dict(name="C2_ragged", kind="workload", workload="C2", kwargs=dict(n_sets=3, n_points=777),
         ragged=[777, 1, 300]),
]

# The reference's adaptive-integration tests, "Single integral"
# (c++/tests/numerical_integration.cpp:13-223): one data set of 150 points,
# p0 = 10, p1 = 1, iteration_limit 4, acceleration 0.9, fit(10.0).  `fits` is
# the list of consecutive fits: (p1 active?, KAT chi2, KAT p0, KAT p1).
INTEGRATION_SCENARIOS = [
    dict(name="single_integral_no_bounds", lines="25-45",
         fits=[(True, 4994.801048103614, 9.345693397983833, 1.086341822060304)]),
    dict(name="single_integral_lower", lines="46-72",
         fits=[(False, 3359.402760955073, 9.638686516377437, 1.0),
               (True, 3359.360525697878, 9.63837358508365, 1.000164288516688)]),
    dict(name="single_integral_lower_nopar", lines="73-95",
         fits=[(False, 3359.374808601714, 9.513801290676248, 1.0)]),
    dict(name="single_integral_upper", lines="96-123",
         fits=[(False, 3359.402760955071, 9.638686516377437, 1.0),
               (True, 3359.360525697879, 9.638373585083652, 1.000164288516688)]),
    dict(name="single_integral_upper_nopar", lines="124-147",
         fits=[(False, 3359.374808601714, 9.513801290676248, 1.0)]),
    dict(name="single_integral_both", lines="148-175",
         fits=[(False, 3359.392136789901, 9.664371097350363, 1.0),
               (True, 3359.360525697834, 9.664108472227593, 1.000124158231295)]),
    dict(name="single_integral_both_lower_fixed", lines="176-199",
         fits=[(False, 96283.63738642586, 4.023936467213234, 1.0)]),
    dict(name="single_integral_both_nopar", lines="200-222",
         fits=[(True, 3359.360587615625, 9.834021674777725, 1.301193106585963)]),
]
for _sc in INTEGRATION_SCENARIOS:
    _sc["kind"] = "integration"


def run_integration(scenario: dict, make_solver, c1_data: dict, n_threads: int = 1) -> list:
    """Runs the consecutive fits of one "Single integral" section; returns one
    record per fit (the second fit warm-starts from the first, as in the
    reference test)."""
    s = make_solver(scenario["name"])
    x = c1_data["x_data_single"].copy()
    y = c1_data["y_data_single"].copy()
    s._scenario_keep = [x, y]
    s.addDataset(x, y)
    s.setPar(0, 10.0, True)
    s.setPar(1, 1.0, True)
    s.settings.iteration_limit = 4
    s.settings.acceleration_threshold = 0.9
    s.settings.n_threads = n_threads
    out = []
    for p1_active, *_ in scenario["fits"]:
        s.setPar(1, 1.0, p1_active)  # numerical_integration.cpp:61,67
        s.fit(10.0)
        out.append(dict(name=scenario["name"], chi2=s.chi2(),
                        pars=[[s.getParValue(0), s.getParValue(1)]], trace=s.trace(),
                        iterations=s.iterations(), dof=s.degreesOfFreedom()))
    return out


# The reference's nested double-integral tests, "Double integral (nested)"
# (c++/tests/numerical_integration.cpp:224-918): 100 points with errors, six
# parameters (7.0, 1.3, 1.2, 2.0, 0.2, 2.1), acceleration 0.9, fit(0.1).
# All sixteen sections (twelve distinct functors: sections that differ only in
# which parameters are active share one).  active = activity of p1..p5 (p0 is
# always active); kat = (chi2, parameters), asserted by the reference at 1e-12.
NESTED_SCENARIOS = [
    dict(name="nested_integral_y1y2x1x2", section="Active bounds: y1 y2 x1 x2", lines="249", active=(1, 1, 1, 1, 1), limit=2,
         kat=(0.2131810550497416, [15.26735468164642, 1.386383105456653, 0.8486391644471797, 1.674240469615365, 0.1885677628244937, 1.941800275111635])),
    dict(name="nested_integral_y1y2x1", section="Active bounds: y1 y2 x1", lines="292", active=(1, 1, 1, 1, 0), limit=2,
         kat=(20529.86214956253, [9.545073737454485, 1.050947728780064, 1.407011447112184, 2.246597745517819, 0.09581638730674913, 2.1])),
    dict(name="nested_integral_y1y2x2", section="Active bounds: y1 y2 x2", lines="336", active=(1, 1, 1, 1, 1), limit=2,
         kat=(8.511262427426729, [31.35420758618348, 1.343236097449233, 0.9880791189004298, 1.915159447508319, 0.6301502301640346, 2.041471780774121])),
    dict(name="nested_integral_y1y2x1x2", section="Active bounds: y1 x1 x2", lines="379", active=(1, 1, 0, 1, 1), limit=2,
         kat=(0.5452442448229686, [14.51912799259439, 1.401803657027402, 0.7703969798462069, 2.0, 0.2243476865643863, 1.911960222088238])),
    dict(name="nested_integral_y2x1x2", section="Active bounds: y2 x1 x2", lines="422", active=(1, 1, 0, 1, 1), limit=2,
         kat=(0.5452442448229419, [14.51912799259429, 1.401803657027403, 0.7703969798462061, 2.0, 0.2243476865643837, 1.911960222088238])),
    dict(name="nested_integral_y1y2", section="Active bounds: y1 y2", lines="465", active=(1, 1, 1, 1, 0), limit=2,
         kat=(31829.01194465925, [8.429293418556341, 1.569188491899031, 1.476297876086944, 2.350856627400455, 0.1445015201991888, 2.1])),
    dict(name="nested_integral_x1x2", section="Active bounds: x1 x2", lines="505", active=(1, 0, 0, 0, 1), limit=2,
         kat=(0.0638207048968614, [15.54318299637472, 1.337653916227864, 1.2, 2.0, 0.2, 2.060422119015556])),
    dict(name="nested_integral_y1y2x2", section="Active bounds: y1 x2", lines="548", active=(1, 1, 0, 1, 1), limit=2,
         kat=(8.310466833011295, [31.05730169163706, 1.337447872754693, 0.9731467848827562, 2.0, 0.6676623753034178, 2.042477682607804])),
    dict(name="nested_integral_y2x1", section="Active bounds: y2 x1", lines="591", active=(0, 0, 1, 0, 1), limit=2,
         kat=(20530.20016213086, [72.09812547421947, 1.3, 1.2, 10.90218525163188, 0.2, 15.56263330043302])),
    dict(name="nested_integral_y1x1", section="Active bounds: y1 x1", lines="628", active=(1, 1, 0, 0, 1), limit=2,
         kat=(0.3708459104616477, [13.42618730322273, 1.40787919245722, 0.7574214524465727, 2.0, 0.2, 1.896290194086714])),
    dict(name="nested_integral_y2x1x2", section="Active bounds: y2 x2", lines="671", active=(1, 1, 0, 0, 1), limit=2,
         kat=(0.3645778424347108, [13.45556877476804, 1.408061308403743, 0.7570256924263207, 2.0, 0.2, 1.895981142726112])),
    dict(name="nested_integral_y1", section="Active bounds: y1", lines="714", active=(1, 0, 0, 0, 0), limit=2,
         kat=(33404.97047824427, [18.74611668457635, 3.127350705902004, 1.2, 2.0, 0.2, 2.1])),
    dict(name="nested_integral_y2", section="Active bounds: y2", lines="754", active=(1, 0, 0, 0, 0), limit=2,
         kat=(33404.97047824427, [18.74611668457635, 3.127350705902005, 1.2, 2.0, 0.2, 2.1])),
    dict(name="nested_integral_x1", section="Active bounds: x1", lines="794", active=(0, 0, 0, 0, 1), limit=2,
         kat=(20529.86874184859, [80.95988477911882, 1.3, 1.2, 2.0, 0.2, 15.63231901313966])),
    dict(name="nested_integral_y1", section="Active bounds: x2", lines="834", active=(0, 0, 0, 0, 1), limit=2,
         kat=(20529.86896231501, [80.95988738910319, 1.3, 1.2, 2.0, 0.2, 15.63232441412172])),
    dict(name="nested_integral_no_active_bounds", section="No active bounds", lines="874", active=(0, 0, 0, 0, 0), limit=1,
         kat=(158.6303014282949, [24.35593003546224, 1.3, 1.2, 2.0, 0.2, 2.1])),
]
for _s in NESTED_SCENARIOS:
    _s["id"] = "nested:" + _s["section"].replace("Active bounds: ", "").replace(" ", "")


def run_nested(scenario: dict, make_solver, c1_data: dict, n_threads: int = 1) -> dict:
    s = make_solver(scenario["name"])
    x = c1_data["x_data_double"].copy()
    y = c1_data["y_data_double"].copy()
    e = c1_data["weights_double"].copy()
    s._scenario_keep = [x, y, e]
    s.addDataset(x, y, e)
    s.setPar(0, 7.0, True)
    for k, (val, act) in enumerate(zip((1.3, 1.2, 2.0, 0.2, 2.1), scenario["active"]), start=1):
        s.setPar(k, val, bool(act))
    s.settings.iteration_limit = scenario["limit"]
    s.settings.acceleration_threshold = 0.9
    s.settings.n_threads = n_threads
    s.fit(0.1)
    return dict(name=scenario["name"], chi2=s.chi2(), pars=[[s.getParValue(k) for k in range(6)]],
                trace=s.trace(), iterations=s.iterations(), dof=s.degreesOfFreedom())


# The reference's direct double-integral tests, "Double integral (direct)"
# (c++/tests/numerical_integration.cpp:930-1440): 100 points without errors,
# seven parameters (7.0, 1.3, 1.2, 2.0, 0.2, 2.1, 1.0), iteration limit 2,
# acceleration 0.9, fit(0.1).  All sixteen sections; `active` = activity of
# p0..p6, kat = (chi2 to 1e-9 absolute, parameters to 1e-7 relative).
DIRECT_SCENARIOS = [
    dict(name="direct_integral_y1y2x1x2", section="Active bounds: y1 y2 x1 x2", lines="951", active=(0, 0, 1, 1, 1, 1, 1),
         kat=(1.654886495874691e-06, [7, 1.3, 2.066882698115843, 2.462337242876448, 0.1286061153388938, 2.370219247402999, 1.537928956329218])),
    dict(name="direct_integral_y1y2x1", section="Active bounds: y1 y2 x1", lines="972", active=(1, 0, 0, 1, 1, 0, 0),
         kat=(4.790523057594158e-09, [9.175204980541729, 1.3, 1.2, 2.516290186341045, 0.1241748448388979, 2.1, 1.0])),
    dict(name="direct_integral_y1y2x2", section="Active bounds: y1 y2 x2", lines="1003", active=(1, 1, 0, 1, 0, 1, 0),
         kat=(8.068219436506581e-09, [8.65075393701988, 1.127842104542012, 1.2, 2.391312159920373, 0.2, 2.307754246034183, 1.0])),
    dict(name="direct_integral_y1x1x2", section="Active bounds: y1 x1 x2", lines="1034", active=(1, 1, 0, 0, 1, 1, 0),
         kat=(7.949911068776061e-09, [8.623217421789654, 1.12999180500759, 1.2, 2.0, 0.143292615316067, 2.304776775635509, 1.0])),
    dict(name="direct_integral_y2x1x2", section="Active bounds: y2 x1 x2", lines="1065", active=(1, 0, 0, 1, 0, 1, 1),
         kat=(0.000213350703610027, [9.671381032914667, 1.3, 1.2, 2.471195537765232, 0.2, 2.436352543805309, 1.282878741340198])),
    dict(name="direct_integral_y1y2", section="Active bounds: y1 y2", lines="1096", active=(1, 1, 0, 1, 0, 0, 0),
         kat=(6.66506150605225e-09, [7.666430772544548, 1.520366990688495, 1.2, 2.149840327725317, 0.2, 2.1, 1.0])),
    dict(name="direct_integral_x1x2", section="Active bounds: x1 x2", lines="1127", active=(1, 0, 0, 0, 1, 1, 0),
         kat=(1.05553491668674e-08, [9.456196153046807, 1.3, 1.2, 2.0, 0.1108266934567069, 2.419211642876849, 1.0])),
    dict(name="direct_integral_y1x2", section="Active bounds: y1 x2", lines="1158", active=(1, 1, 0, 0, 0, 1, 0),
         kat=(1.837877829573166e-08, [9.13367142357661, 1.077434702465759, 1.2, 2.0, 0.2, 2.369246887944458, 1.0])),
    dict(name="direct_integral_y2x1", section="Active bounds: y2 x1", lines="1189", active=(1, 0, 0, 1, 0, 0, 1),
         kat=(3.85575335670611e-09, [9.161296077178266, 1.3, 1.2, 2.513226918313678, 0.2, 2.1, 0.6086751941736143])),
    dict(name="direct_integral_y1x1", section="Active bounds: y1 x1", lines="1220", active=(1, 0, 0, 0, 1, 0, 0),
         kat=(4.556742521509683e-09, [9.972875346712668, 1.3, 1.2, 2.0, 0.09633003605472064, 2.1, 1.0])),
    dict(name="direct_integral_y2x2", section="Active bounds: y2 x2", lines="1251", active=(1, 0, 0, 1, 0, 1, 0),
         kat=(5.373563892617068e-08, [9.405485170085401, 1.3, 1.2, 2.5661598112606, 0.2, 2.403368518336621, 1.0])),
    dict(name="direct_integral_y1", section="Active bounds: y1", lines="1282", active=(1, 1, 0, 0, 0, 0, 0),
         kat=(1.443756776956618e-07, [8.13832625826087, 1.657624293024702, 1.2, 2.0, 0.2, 2.1, 1.0])),
    dict(name="direct_integral_y2", section="Active bounds: y2", lines="1313", active=(0, 0, 0, 1, 0, 0, 0),
         kat=(0.3053680170120716, [7, 1.3, 1.2, 6.663707134981233, 0.2, 2.1, 1.0])),
    dict(name="direct_integral_x1", section="Active bounds: x1", lines="1345", active=(0, 0, 0, 0, 1, 0, 0),
         kat=(7.119221262116694e-07, [7, 1.3, 1.2, 2.0, 0.02430156976447609, 2.1, 1.0])),
    dict(name="direct_integral_x2", section="Active bounds: x2", lines="1377", active=(0, 0, 0, 0, 0, 1, 0),
         kat=(2.851428619095947e-06, [7, 1.3, 1.2, 2.0, 0.2, 3.034543683583202, 1.0])),
    dict(name="direct_integral_y1y2x1x2", section="No active bounds", lines="1409", active=(1, 0, 0, 0, 0, 0, 0),
         kat=(4.090863893678671e-05, [16.70423680614829, 1.3, 1.2, 2.0, 0.2, 2.1, 1.0])),
]
for _s in DIRECT_SCENARIOS:
    _s["id"] = "direct:" + _s["section"].replace("Active bounds: ", "").replace(" ", "")


def run_direct(scenario: dict, make_solver, c1_data: dict, n_threads: int = 1, limit: int = 2,
               keep: bool = False) -> dict:
    s = make_solver(scenario["name"])
    x = c1_data["x_data_double"].copy()
    y = c1_data["y_data_double_direct"].copy()
    s._scenario_keep = [x, y]
    s.addDataset(x, y)
    for k, (val, act) in enumerate(zip((7.0, 1.3, 1.2, 2.0, 0.2, 2.1, 1.0), scenario["active"])):
        s.setPar(k, val, bool(act))
    s.settings.iteration_limit = limit
    s.settings.acceleration_threshold = 0.9
    s.settings.n_threads = n_threads
    s.fit(0.1)
    rec = dict(name=scenario["id"], chi2=s.chi2(), pars=[[s.getParValue(k) for k in range(7)]],
               trace=s.trace(), iterations=s.iterations(), dof=s.degreesOfFreedom())
    if keep:
        rec["solver"] = s
    return rec


ALL_SCENARIOS = C1_SCENARIOS + SCALED_SCENARIOS


def build(scenario: dict, make_solver, c1_data: dict):
    """Instantiate `scenario` on a solver produced by make_solver(model)."""
    keep = []
    if scenario["kind"] == "c1":
        d = c1_data["fix_d"]
        s = make_solver("exponential")
        n1, n2 = scenario["n_points"]
        x1, y1 = c1_data["x_data_1"][:n1].copy(), c1_data["y_data_1"][:n1].copy()
        x2, y2 = c1_data["x_data_2"][:n2].copy(), c1_data["y_data_2"][:n2].copy()
        keep += [x1, y1, x2, y2]
        s.addDataset(x1, y1)
        s.addDataset(x2, y2)
        a = scenario["active"]
        i00, i01 = scenario["i0"]
        s.setPar(0, d[i00], bool(a[0]), 0)
        s.setPar(2, d[BGR_0], bool(a[1]), 0)
        s.setPar(0, d[i01], bool(a[2]), 1)
        s.setPar(2, d[BGR_1], bool(a[3]), 1)
        s.setPar(1, d[TAU], bool(a[4]))
        for k, v in scenario["settings"].items():
            setattr(s.settings, k, v)
        lam = scenario["lambda0"]
        n_sets, n_pars = 2, 3
    else:
        w = W.WORKLOADS[scenario["workload"]](**scenario["kwargs"])
        s = make_solver(w.model)
        if "ragged" in scenario:
            # unequal data-set sizes (incl. a single-point set): edge case of
            # the tile / slot bookkeeping
            for i, n in enumerate(scenario["ragged"]):
                xs, ys, es = w.slice(i)
                s.addDataset(xs[:n], ys[:n], es[:n])
            for k in w.shared:
                s.setPar(k, float(w.start[0, k]), True)
            for key, val in w.settings.items():
                setattr(s.settings, key, val)
        else:
            w.setup(s)
        keep.append(w)
        lam = w.lambda0
        n_sets, n_pars = w.n_sets, w.n_pars
    s._scenario_keep = keep
    return s, lam, n_sets, n_pars


def run(scenario: dict, make_solver, c1_data: dict, n_threads: int = 1) -> dict:
    s, lam, n_sets, n_pars = build(scenario, make_solver, c1_data)
    s.settings.n_threads = n_threads
    s.fit(lam)
    return dict(
        name=scenario["name"],
        chi2=s.chi2(),
        pars=[[s.getParValue(k, i) for k in range(n_pars)] for i in range(n_sets)],
        trace=s.trace(),
        iterations=s.iterations(),
        dof=s.degreesOfFreedom(),
    ), s


# One record per compare_to_golden() call of this test session; printed by
# tests/conftest.py at the end of the run and checked by
# tests/test_zz_parity_budget.py (how many comparisons NEEDED the relaxed
# parameter tolerance).
PARITY_LOG: list = []
STRICT_PAR_TOL = 1e-10   # north star: parameters and chi2 to 1e-10 relative
RELAXED_PAR_TOL = 1e-7   # only after chi2 has converged to rounding noise


class ParityResult(int):
    """Number of trials compared (an int, for the callers' `n >= 3` checks) +
    the record of what was compared and at which tolerance."""
    record: dict


def _par_error(gp, wp):
    """Largest |got - want| relative to |want|, with an absolute floor of 1e-11
    of the largest magnitude the parameter takes over the data sets: a local
    parameter that happens to sit near zero in one data set (e.g. a background
    slope of 2e-3) has no meaningful relative error."""
    floor = 1e-11 * np.max(np.abs(wp), axis=0, keepdims=True)
    excess = np.abs(gp - wp) - floor
    return float(np.max(np.maximum(excess, 0.0) / np.maximum(np.abs(wp), 1e-300)))


def compare_to_golden(got: dict, want: dict, rtol: float, converged_rel: float = 1e-12,
                      label: str = ""):
    """chi2 and the accept/reject sequence of `got` against `want`:

    * every trial up to and including the first one whose chi2 change is
      <= `converged_rel` relative must have the same decision, the same lambda
      and new_chi2 within rtol.  Beyond that point the reference's own
      decisions are rounding noise (its test comment,
      c++/tests/lm_solver.cpp:396-398); if no such trial exists the whole trace
      and the iteration count must be identical;
    * final chi2 within max(rtol, 1e-12);
    * parameters within max(rtol, 1e-10) (STRICT, the north star's tolerance).
      Only if the trace comparison stopped at a converged trial AND the strict
      check fails, 1e-7 is accepted (the remaining iterations move the
      parameters by rounding noise times the conditioning) -- and the record
      says so: `needed_relaxed`.  tests/test_zz_parity_budget.py bounds how
      many comparisons of a session may need that.

    Returns the number of trials compared (ParityResult, an int with a
    `.record`)."""
    tg, tw = got["trace"], want["trace"]
    n_cmp = 0
    stopped_at_convergence = False
    for a, b in zip(tg, tw):
        assert a["decision"] == b["decision"], (got["name"], n_cmp, a, b)
        assert a["lambda_"] == b["lambda_"], (got["name"], n_cmp, a, b)
        np.testing.assert_allclose(a["new_chi2"], b["new_chi2"], rtol=rtol)
        n_cmp += 1
        if abs(b["new_chi2"] - b["old_chi2"]) <= converged_rel * abs(b["old_chi2"]):
            stopped_at_convergence = n_cmp < len(tw)
            break
    else:
        assert len(tg) == len(tw), (got["name"], len(tg), len(tw))
        assert got["iterations"] == want["iterations"]
    # how far the two traces agree in fact (decisions and lambda), for the record
    n_same = 0
    for a, b in zip(tg, tw):
        if a["decision"] != b["decision"] or a["lambda_"] != b["lambda_"]:
            break
        n_same += 1
    identical_to_end = n_same == len(tw) == len(tg) and got["iterations"] == want["iterations"]
    np.testing.assert_allclose(got["chi2"], want["chi2"], rtol=max(rtol, 1e-12))
    gp, wp = np.array(got["pars"]), np.array(want["pars"])
    worst = _par_error(gp, wp)
    strict = max(rtol, STRICT_PAR_TOL)
    needed_relaxed = False
    if worst > strict:
        # the relaxed tolerance exists only for the post-convergence tail
        assert stopped_at_convergence and not identical_to_end, (
            got["name"], "parameters differ by", worst, "with an identical trace: tolerance", strict)
        assert worst <= max(rtol, RELAXED_PAR_TOL), (got["name"], worst)
        needed_relaxed = True
    assert got["dof"] == want["dof"]
    chi2_rel = abs(got["chi2"] - want["chi2"]) / max(abs(want["chi2"]), 1e-300)
    result = ParityResult(n_cmp)
    result.record = dict(
        name=got["name"], label=label, trials=len(tw), trials_compared=n_cmp,
        trials_identical=n_same, trace_identical_to_end=bool(identical_to_end),
        stopped_at_converged_trial=bool(stopped_at_convergence),
        worst_parameter_error=worst, chi2_error=chi2_rel,
        parameter_tolerance=RELAXED_PAR_TOL if needed_relaxed else strict,
        needed_relaxed=needed_relaxed)
    PARITY_LOG.append(result.record)
    return result
