# dump_reference.jl — for anyone with Julia >= 1.12 and LevelSetMethods.jl v0.2.0:
# run the REAL reference on the parity configurations of tests/helpers.py and dump phi after N RK3 steps as raw
# little-endian Float64 (column-major), so the CPU oracle (oracle/) can be pinned at the array level.
# Not runnable in this repository's images (no Julia); see DESIGN.md §4 "residual risk".
#
#   julia --project tools/dump_reference.jl outdir
using LevelSetMethods, StaticArrays, LinearAlgebra
const LSM = LevelSetMethods

function run_case(name, grid, ϕ0, terms, bc, nsteps; integrator = RK3())
    ϕ = MeshField(ϕ0, grid)
    eq = LevelSetEquation(; terms, ic = ϕ, bc, integrator)
    dt0 = LSM.cfl(integrator) * LSM.compute_cfl(eq.terms, current_state(eq), 0.0)
    integrate!(eq, dt0 * nsteps * (1 - 1e-12))
    open(joinpath(ARGS[1], "$name.f64"), "w") do io
        write(io, values(current_state(eq)))
    end
    println(name, ": t = ", current_time(eq), "  min/max = ", extrema(values(current_state(eq))))
end

mkpath(ARGS[1])
# C1: 2-D circle under rigid rotation, 128^2, periodic
g = CartesianGrid((-1, -1), (1, 1), (128, 128))
u = MeshField(x -> SVector(-x[2], x[1]), g)
run_case("C1_128", g, x -> hypot(x[1] - 0.3, x[2]) - 0.4, (AdvectionTerm(u),), PeriodicBC(), 100)
# C4: Eikonal reinitialisation of a perturbed sphere, 48^3, Neumann
g = CartesianGrid((-1, -1, -1), (1, 1, 1), (48, 48, 48))
f4 = x -> (norm(x) - 0.5) * (1 + 0.4 * sin(3π * x[1]) * sin(3π * x[2]) * sin(3π * x[3]))
run_case("C4_48", g, f4, (EikonalReinitializationTerm(MeshField(f4, g)),), NeumannBC(), 50)
# C5: NormalMotion(v = 0.2 field) + Advection(u = (-y, x, 0) field), 48^3, Neumann
v = MeshField(x -> 0.2, g); u3 = MeshField(x -> SVector(-x[2], x[1], 0.0), g)
run_case("C5_48", g, x -> norm(x .- SVector(0.3, 0.0, 0.0)) - 0.4, (NormalMotionTerm(v), AdvectionTerm(u3)), NeumannBC(), 100)

# ---- "next" rows (SURVEY.md §8f): scalars and arrays for volume / perimeter, set operations, velocity extension ----
function dump(name, A)
    open(io -> write(io, A), joinpath(ARGS[1], "$name.f64"), "w")
end
g2 = CartesianGrid((-1.0, -1.0), (1.0, 1.0), (81, 61))
a = MeshField(x -> hypot(x[1] - 0.2, x[2]) - 0.5, g2)
b = MeshField(x -> max(abs(x[1] + 0.1), abs(x[2])) - 0.4, g2)
println("volume(a) = ", repr(LSM.volume(a)), "  perimeter(a) = ", repr(LSM.perimeter(a)))
dump("csg_union", values(union(a, b))); dump("csg_intersect", values(intersect(a, b)))
dump("csg_setdiff", values(setdiff(a, b))); dump("csg_complement", values(LSM.complement(a)))
# extend_along_normals! on the grid of test/test-velocityextension.jl (default band mask, 12 iterations)
ϕe = MeshField(x -> (hypot(x[1], x[2]) - 0.5) * (1 + 0.2 * x[1]), g2)
F = [sin(3 * x) + y^2 for x in LSM.grid1d(g2, 1), y in LSM.grid1d(g2, 2)]
extend_along_normals!(F, ϕe; nb_iters = 12)
dump("extend_81x61", F)
