"""Known-answer values transcribed from the reference's own tests, doctests and docs.

Every entry cites the file:line under /root/reference it was read from.  Inputs that the
values depend on are the committed fixtures in this directory (see make_golden.py).
"""
import numpy as np

# --- NAMD: LJ energy of 9 999 Ne atoms, cutoff 10 --------------------------------------
# lj_NE: u += eps*((sig/d)^12 - 2(sig/d)^6), eps=0.0441795, sig=2*1.64009
# (test/applications/namd/compare_with_namd.jl:7-13)
NAMD_EPS = 0.0441795
NAMD_SIG = 2 * 1.64009
NAMD_C12 = NAMD_EPS * NAMD_SIG ** 12          # same potential in c12/d2^6 - c6/d2^3 form
NAMD_C6 = 2 * NAMD_EPS * NAMD_SIG ** 6
NAMD_CUTOFF = 10.0
_T45 = np.array([[70.7107, 0.0, 0.0], [35.3553, 61.2372, 0.0], [35.3553, 20.4124, 57.735]]).T
NAMD_CASES = {
    # frame: (unitcell, golden energy)           compare_with_namd.jl line
    "o1": ([50.0, 50.0, 50.0], 32230.01699504111),                                             # :77-79
    "o2": ([80.0, 70.0, 50.0], 1093.7225407797744),                                            # :81-83
    "o3": (np.array([[50.0, 0.0, 50.0], [50.0, 50.0, 0.0], [0.0, 50.0, 50.0]]), 1724.3195067566828),  # :87-92
    "o4": (_T45, 1754.0802503953591),                                                          # :94-99
    "o5": (_T45, 1765.1389457850137),                                                          # :102-111
    "o6": ([80.0, 80.0, 80.0], -158.04751357760088),                                           # :113-115
    "t1": (np.array([[80.0, 0.0, 30.0], [30.0, 80.0, 0.0], [0.0, 40.0, 80.0]]), -116.53213607052128),  # :119-124
    "t2": (np.array([[50.0, 0.0, 0.0], [50.0, 50.0, 0.0], [0.0, 50.0, 50.0]]), 32096.48839031735),     # :127-132
}
NAMD_LCELLS = (1, 3)  # the same goldens are repeated with lcell = 3 (:168-224)

# --- argon_pdb_file doctests (100 Ar atoms, test/applications/gromacs/argon/cubic.pdb) --
ARGON_UNITCELL = [21.0, 21.0, 21.0]
ARGON_CUTOFF = 8.0
ARGON_SUM_D2 = 43774.54367600001               # src/API/ParticleSystem.jl:110-118
ARGON_SUM_D2_CROSS = 21886.196785000004        # x = atoms 1:50, y = 51:100; src/API/ParticleSystem.jl:125-139
ARGON_NL_NONPERIODIC = (857, (1, 20, 3.163779526466901))    # src/API/neighborlist.jl:265-276
ARGON_NL_PERIODIC = (1143, (1, 7, 3.3638756414119397))      # src/API/neighborlist.jl:280-292
ARGON_NL_CROSS = (439, (1, 11, 4.08865164675529))           # src/API/neighborlist.jl:296-310 (x: index<=50, y: index>50, non-periodic)
ARGON_MIN_DIST = 2.1991993997816563            # src/API/parallel_custom.jl:183-192
ARGON_SUM_INV_D = 207.37593043370862           # docs/src/ParticleSystem/single_set_compound.md:108-109
ARGON_FORCES_INV_D = {                         # f_i = sum_j (x_j - x_i)/d^3 ; single_set_compound.md:111-117
    0: [0.02649383330735732, 0.18454277989323772, -0.012253902366284958],
    1: [0.07782602581235692, 0.27910822337402613, 0.21926615329195248],
    98: [0.11307234751448932, 0.006353545239676281, -0.05955687310348303],
    99: [-0.031012009183076745, 0.03543655648545698, 0.03184912163097636],
}

# --- grid KATs -------------------------------------------------------------------------
GRID_KATS = [
    # (unitcell, cutoff, lcell, nc, cell_size)
    ([10.0, 10.0, 10.0], 1.0, 1, [13, 13, 13], [1.0, 1.0, 1.0]),          # test/API/test_show.jl:11-20
    ([10.0, 10.0, 10.0], 0.1, 1, [103, 103, 103], [0.1, 0.1, 0.1]),       # test/API/test_show.jl:55-64
    ([120.0, 150.0, 100.0], 10.0, 1, [15, 18, 13], [10.0, 10.0, 10.0]),   # src/internals/Box.jl:317-327
    (np.array([[100.0, 50.0, 0.0], [0.0, 120.0, 0.0], [0.0, 0.0, 130.0]]), 10.0, 1, [20, 13, 16], [10.0, 10.0, 10.0]),  # Box.jl:156-170
]
# 100 points anywhere in [0,1)^3 inside Box([10,10,10],1): 1 cell with real particles, 800 particles incl. images
SHOW_CELLLIST_KAT = dict(n=100, unitcell=[10.0, 10.0, 10.0], cutoff=1.0, n_cells_real=1, n_particles=800)  # test/API/test_show.jl:22-29
# get_computing_box for the unit cube, cutoff 0.1 -> ([-0.1]^3, [1.1]^3)   test/API/ParticleSystem.jl:228-234
COMPUTING_BOX_KAT = dict(unitcell=[1.0, 1.0, 1.0], cutoff=0.1, lo=[-0.1] * 3, hi=[1.1] * 3)

# --- wrapping KAT: wrap_to_first([15,13],[10 0;0 10]) = [5.0, 3.0000000000000004] -------
WRAP_KAT = ([15.0, 13.0], [10.0, 10.0], [5.0, 3.0000000000000004])        # src/internals/CellOperations.jl:79-88

# --- boundary KATs: test/API/neighborlists.jl:183-192, :223-260 ; test/internals/tests.jl:428-437, :482-486
nf = lambda v: float(np.nextafter(v, np.inf))
pf = lambda v: float(np.nextafter(v, -np.inf))
BOUNDARY_KATS = [
    # (x, cutoff, unitcell, expected number of pairs)
    ([[0.0, 0.0, 1.0], [0.0, 0.0, 10.0], [0.0, 0.0, 7.0]], 2.0, None, 0),
    ([[0.0, 0.0, 1.0], [0.0, 0.0, 10.0]], 2.0, None, 0),
    ([[0.0, 1.0], [0.0, 10.0]], 2.0, None, 0),
    ([[0.0, 1.0]], 2.0, None, 0),
    ([[0.0, 0.0]], 2.0, None, 0),
    ([[0.0, 0.0, 0.0]], 2.0, None, 0),
    ([[0.0, 0.0]], 1.0, [2.0 + nf(1.0), 2.0 + nf(1.0)], 0),
    ([[0.0, 0.0], [0.0, 1.0]], 1.0, [2.0 + nf(1.0), 2.0 + nf(1.0)], 1),           # d == cutoff exactly: (1,2,1.0)
    ([[0.0, 0.0], [0.0, 1.0]], pf(1.0), [2.0, 2.0], 0),
    ([[0.0, 0.0], [0.0 + nf(1.0), 1.0 + nf(1.0)]], pf(1.0), [2.0, 2.0], 1),       # d = 0.9999999999999998
    ([[0.0, 0.0], [-1.0, 0.0]], 5.0, [14.01, 14.02], 1),
    ([[0.0, 0.0], [nf(0.1), 0.0]], 0.1, [1.0, 1.0], 0),
    ([[0.0, 0.0], [pf(0.9), 0.0]], 0.1, [1.0, 1.0], 0),
    ([[0.0, 0.0], [-0.1, 0.0]], 0.1, [1.0, 1.0], 1),
    ([[0.0, 0.0], [0.1, 0.0]], 0.1, [1.0, 1.0], 1),
    ([[0.0, 0.0], [0.9, 0.0]], 0.1, [1.0, 1.0], 1),
    ([[1.0, 2.0], [3.0, 4.0]], 3.0, None, 1),
    # sph case (https://github.com/m3g/CellListMap.jl/issues/95): exactly one pair
    ([[0.0, 2.52], [0.02, 2.56], [3.98, 2.96], [4.0, 0.26], [4.0, 2.5]], 0.06788225099390856, None, 1),
    # exactly-once counting across PBC, test/internals/tests.jl:494-507
    ([[0.5, 5.0, 5.0], [9.6, 5.0, 5.0]], 2.0, [10.0, 10.0, 10.0], 1),
]
BOUNDARY_D_KAT = 0.9999999999999998   # distance reported in the nextfloat case above (neighborlists.jl:192)
# Float32 pathological sets (bug 84): lists must have unique pairs  (neighborlists.jl:195-218)
BUG84_3D = [[0.0, 0.0, 0.0], [0.154, 1.136, -1.827], [-1.16, 1.868, 4.519], [-0.089, 2.07, 4.463], [0.462, -0.512, 5.473]]
BUG84_2D = [[0.0, 0.0], [0.0, -2.0], [-0.1, 5.0], [0.0, 5.5]]
# few particles (neighborlists.jl:250-256): cross x=[1,1,1], y=[[1.05,1,1],[0,0,0]] -> (1,1,0.05); self z -> (1,2,0.05)
FEW_CROSS = ([[1.0, 1.0, 1.0]], [[1.05, 1.0, 1.0], [0.0, 0.0, 0.0]], 0.1, (1, 1, 0.05))
FEW_SELF = ([[1.0, 1.0, 1.0], [1.05, 1.0, 1.0], [0.0, 0.0, 0.0]], 0.1, (1, 2, 0.05))
# triclinic exactly-once set (test/internals/tests.jl:509-528): cell-list count must equal naive count
TRICLINIC_ONCE = dict(
    unitcell=np.array([[10.0, 3.0, 0.0], [0.0, 10.0, 2.0], [0.0, 0.0, 10.0]]), cutoff=2.5,
    x=[[0.5, 0.5, 0.5], [2.0, 1.0, 1.0], [9.5, 9.5, 9.5], [8.5, 9.0, 9.0]])
# pathological 2-D cells, cutoff 0.2 (test/internals/tests.jl:454-476)
_l = np.sqrt(2) / 2
PATHOLOGICAL_2D_CELLS = [
    np.array([[1.0, 0.0], [0.0, 1.0]]), np.array([[_l, 0.0], [_l, 1.0]]), np.array([[1.1, 0.0], [0.0, 1.0]]),
    np.array([[1.2, 0.0], [0.0, 1.0]]), np.array([[1.0, 0.0], [0.0, 1.1]]), np.array([[1.0, 0.0], [0.0, 1.2]]),
    np.array([[1.0, 0.2], [0.0, 1.2]]), np.array([[1.0, 0.2], [0.2, 1.2]]), np.array([[1.2, 0.2], [0.2, 1.2]]),
]
# test_pathological matrices (test/modules/Testing.jl:572-585) incl. negative entries are generated in the tests.


# cell_limits(align_cell(m)) (test/internals/CellOperations.jl:159-194): bounding box of the aligned cell's vertices.
# (matrix with COLUMNS = lattice vectors as in the reference, expected lo, expected hi, exact?)
CELL_LIMITS_KATS = [
    ([[10.0, 5.0], [5.0, 10.0]], [0.0, 0.0], [20.12461179749811, 6.708203932499369], False),          # :168-171
    ([[10.0, 5.0], [0.0, 10.0]], [0.0, -8.94427190999916], [15.652475842498529, 0.0], False),          # :176-179
    ([[1.0, 0.0, 0.0], [0.0, 2.0, 0.0], [0.0, 0.0, 1.0]], [0.0, -1.0, 0.0], [2.0, 0.0, 1.0], True),    # :184-188
    ([[1.0, 0.0, 0.0], [0.0, 2.0, 0.0], [0.0, 0.0, 3.0]], [0.0, 0.0, -1.0], [3.0, 2.0, 0.0], True),    # :190-194
]

# align_cell (test/internals/CellOperations.jl:113-128): (m, aligned m, rotation R) with l = sqrt(2)/2; COLUMNS = lattice vectors
_L = 2.0 ** 0.5 / 2.0
ALIGN_CELL_KATS = [
    ([[_L, 0.0], [_L, 1.0]], [[1.0, _L], [0.0, _L]], [[_L, _L], [-_L, _L]]),          # :117-120
    ([[-_L, 0.0], [_L, 1.0]], [[1.0, _L], [0.0, -_L]], [[-_L, _L], [-_L, -_L]]),       # :122-125
]
# wrap_relative_to (test/internals/CellOperations.jl:7-26): (x, y, wrap(x rel. y), wrap(y rel. x)), cell sides +-10
WRAP_RELATIVE_KATS = [
    ([15.0, 13.0], [4.0, 2.0], [5.0, 3.0], [14.0, 12.0]),
    ([-7.0, -6.0], [1.0, 2.0], [3.0, 4.0], [-9.0, -8.0]),
]
