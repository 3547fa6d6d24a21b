# TorJCuda.jl — the ccall shim of INTEGRATION.md as a file a maintainer can `include` after src/plasma.jl.
# NOT EXECUTED in the build environment of this repository (Julia is not installed there); the same C ABI calls, in the
# same order, are exercised by the Python mirror torj_jl_b200/ and by tests/c_abi/abi_smoke.c.
# It replaces the bodies of abs_Al_init (reference src/absorption.jl:1-7), make_ray (src/solve.jl:135-181) and
# make_beam (src/solve.jl:209-242); Plasma(...) and launch_peripheral_rays stay as they are.

const libtorj = joinpath(@__DIR__, "..", "deps", "libtorj_cuda.so")

struct TorjOptions          # torj_options, field for field
    scheme::Int32; n_segments::Int32
    dtmax::Float64; abstol::Float64; reltol::Float64; psi_stop::Float64; p_stop::Float64; te_min::Float64
    max_harmonic::Int32; max_steps_per_segment::Int32
    alpha_floor::Float64
    schedule::Int32; reserved_::Int32
end
struct TorjGrid; nR::Int32; nZ::Int32; R_first::Float64; R_last::Float64; Z_first::Float64; Z_last::Float64; end
struct TorjCounters; n_acc::Int64; n_rej::Int64; n_rhs::Int64; n_alpha::Int64; n_harm::Int64; n_rays_ok::Int64; n_harm_pruned::Int64; n_alpha_skipped::Int64; end

torj_check(rc) = rc == 0 || error(unsafe_string(ccall((:torj_last_error, libtorj), Cstring, ())))

const _ctx = Ref{Ptr{Cvoid}}(C_NULL)
function torj_ctx()
    if _ctx[] == C_NULL
        @assert ccall((:torj_abi_version, libtorj), Cint, ()) == 2   # the struct layouts above are TORJ_ABI_VERSION 2
        torj_check(ccall((:torj_ctx_create, libtorj), Cint, (Cint, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), 0, C_NULL, _ctx))
    end
    _ctx[]
end

# abs_Al_init(N)  — src/absorption.jl:1-7 keeps filling the Julia globals and additionally uploads them
function abs_Al_init(N_absz::Int)
    global _int_weights, _int_absz
    _int_absz, _int_weights = gausslegendre(N_absz)
    torj_check(ccall((:torj_abs_init, libtorj), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}),
                     torj_ctx(), N_absz, _int_absz, _int_weights))
end

# Device tables from a Plasma. Two routes:
#  (a) robust, needs no Interpolations.jl internals: keep the constructor's raw arrays in the struct and call
#      torj_plasma_create_from_data (1-D resampling/log/prefilter on the host, 2-D prefilters + packing on the GPU);
#  (b) hand over the already prefiltered padded coefficient matrices of the six 2-D splines and the 1-D V(ψ) spline:
# (Extrapolation -> ScaledInterpolation -> BSplineInterpolation; if those internals move, evaluate each spline at
#  its grid nodes and call torj_bspline_prefilter_2d instead — the natural cubic spline is unique.)
coefs(spl) = Matrix{Float64}(parent(spl.itp.itp.coefs))
function torj_plasma(p::Plasma)
    g = TorjGrid(length(p.R_coords), length(p.Z_coords), p.R_coords[1], p.R_coords[end], p.Z_coords[1], p.Z_coords[end])
    vr = p.volume_psi_spline.itp.ranges[1]
    vc = Vector{Float64}(parent(p.volume_psi_spline.itp.itp.coefs))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    torj_check(ccall((:torj_plasma_create, libtorj), Cint,
        (Ptr{Cvoid}, Ref{TorjGrid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Ptr{Float64}, Int32, Float64, Float64, Float64, Ref{Ptr{Cvoid}}),
        torj_ctx(), g, coefs(p.psi_norm_spline), coefs(p.ne_spline), coefs(p.Te_spline), coefs(p.Br_spline),
        coefs(p.Bz_spline), coefs(p.Bϕ_spline), vc, length(vr), first(vr), step(vr), p.psi_prof_max, h))
    h[]     # attach a finalizer calling torj_plasma_destroy
end

function torj_trace(hp, pos::Matrix{Float64}, dir::Matrix{Float64}, w::Vector{Float64}, f, mode, s_max, psi;
                    traj::UnitRange=1:0, max_pts=0)
    n = length(w); npsi = length(psi); nt = length(traj)
    opt = Ref{TorjOptions}(); ccall((:torj_options_default, libtorj), Cvoid, (Ref{TorjOptions},), opt)
    dP_dV = zeros(npsi); dep = Ref(0.0); Pf = zeros(n); Pd = zeros(n); npts = zeros(Int32, n); st = zeros(Int32, n)
    ts = zeros(max_pts, nt); txyz = zeros(max_pts, 3, nt); tP = zeros(max_pts, nt); tdP = zeros(max_pts, nt); tprof = zeros(npsi, nt)
    cnt = Ref{TorjCounters}()
    torj_check(ccall((:torj_trace, libtorj), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ref{TorjOptions}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Float64}, Ref{Int32}, Int32,
         Float64, Int32, Ptr{Float64}, Int32, Ptr{Int32}, Ptr{Float64}, Ref{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32},
         Int64, Int64, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{TorjCounters}),
        torj_ctx(), hp, opt, n, pos, dir, w, Float64(f), Int32(mode), 0, s_max, npsi, psi, 1, C_NULL, dP_dV, dep, Pf, Pd, npts, st,
        first(traj) - 1, nt, max_pts, ts, txyz, tP, tdP, tprof, cnt))
    (; dP_dV, deposited = dep[], Pf, Pd, npts, st, ts, txyz, tP, tdP, tprof)
end

# make_ray — src/solve.jl:135-181, same signature and return tuple
function make_ray(plasma::Plasma, x0, N_vacuum, f, mode::Integer, s_max::Float64, psi_dP_dV)
    r = torj_trace(torj_plasma(plasma), reshape(collect(x0), 1, 3), reshape(collect(N_vacuum), 1, 3), [1.0], f, mode, s_max,
                   psi_dP_dV; traj=1:1, max_pts=2 + 100 * (ceil(Int, s_max / 100 / 1e-4) + 8))
    r.st[1] == 0 || throw(AssertionError("ray status $(r.st[1])"))     # src/solve.jl:32,138,141
    n = r.npts[1]
    return r.ts[1:n, 1], [r.txyz[i, :, 1] for i in 1:n], r.tP[1:n, 1], r.tprof[:, 1], r.Pd[1]
end

# make_beam — src/solve.jl:209-242: the Dagger.@spawn loop and the weighted reduction become one call
function make_beam(plasma::Plasma, r, phi, z, tor, pol, spot, invRc, f, mode::Integer, s_max::Float64, psi_dP_dV; kwargs...)
    N0 = collect(IMAS.pol_tor_angles_2_vector(pol, tor)); x0 = [r * cos(phi), r * sin(phi), z]
    pos, dir, w = launch_peripheral_rays(x0, N0, spot, invRc, f; kwargs...)
    n = length(w)
    q = torj_trace(torj_plasma(plasma), pos, dir, w, f, mode, s_max, psi_dP_dV; traj=1:n,
                   max_pts=2 + 100 * (ceil(Int, s_max / 100 / 1e-4) + 8))
    all(q.st .== 0) || throw(AssertionError("ray status $(q.st)"))
    arc = [q.ts[1:q.npts[i], i] for i in 1:n]
    traj = [[q.txyz[k, :, i] for k in 1:q.npts[i]] for i in 1:n]
    pw = [q.tP[1:q.npts[i], i] for i in 1:n]
    return arc, traj, pw, q.dP_dV, q.deposited, w
end
