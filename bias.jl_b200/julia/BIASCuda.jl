#=
BIASCuda.jl — the reference-side binding: routes BIAS.jl's sampler entry points to libbias_cuda.

Written in the reference's own dialect (Julia 0.4, REQUIRE:1).  For Julia >= 0.7 replace
`immutable` -> `struct`, `type` -> `mutable struct`, `Ptr{Void}` -> `Ptr{Cvoid}`,
`Array(T, n)` -> `Vector{T}(undef, n)`.  Julia is not available in the build image, so this file is
untested there; the same C ABI is exercised by the ctypes binding (bias.jl_b200/_lib.py) that the
test-suite drives.

Usage inside BIAS.jl (src/BIAS.jl, after the includes):

    include("BIASCuda.jl")          # defines BIASCuda
    # collapsed_gibbs_sampler!(lda, xx, zz, ...) now dispatches to the GPU when
    # BIASCuda.enabled() is true; the Julia API (model constructors, conjugate types, zz mutated
    # in place, HDP's return tuple) is unchanged.
=#
module BIASCuda

const LIB = get(ENV, "BIAS_CUDA_LIB", "libbias_cuda")

# mirrors include/bias_cuda.h
const BIAS_MD, BIAS_G1G1 = Int32(0), Int32(1)
const BIAS_INT, BIAS_SENT, BIAS_F64 = Int32(0), Int32(1), Int32(2)

immutable Conjugate            # bias_conjugate
    kind::Int32
    datum::Int32
    dd::Int64
    aa::Float64
    m0::Float64
    v0::Float64
    vv::Float64
end

immutable Data                 # bias_data
    n_items::Int64
    word::Ptr{Int32}
    x::Ptr{Float64}
    sent_ptr::Ptr{Int64}
    sent_word::Ptr{Int32}
    sent_count::Ptr{Int32}
end

immutable SweepStats           # bias_sweep_stats
    log_likelihood::Float64
    n_moved::Int64
    KK::Int32
    reserved::Int32
    gg::Float64
    aa::Float64
end

type RcrpParams                # bias_rcrp_params (mutable: KK comes back updated)
    KK::Int32
    Kmax::Int32
    a1::Float64
    a2::Float64
    sample_hyperparam::Int32
    n_internals::Int32
end

immutable RcrfParams           # bias_rcrf_params
    KK::Int32
    Kmax::Int32
    g1::Float64
    g2::Float64
    a1::Float64
    a2::Float64
    sample_hyperparam::Int32
    n_internals::Int32
    join_tables::Int32
end

type HdpParams                 # bias_hdp_params (mutable: KK, gg, aa come back updated)
    KK::Int32
    Kmax::Int32
    gg::Float64
    g1::Float64
    g2::Float64
    aa::Float64
    a1::Float64
    a2::Float64
    sample_hyperparam::Int32
    n_internals::Int32
end

last_error() = bytestring(ccall((:bias_last_error, LIB), Ptr{UInt8}, ()))

function check(rc::Cint)
    rc == 0 && return
    msg = last_error()
    rc == 1 && throw(ArgumentError(msg))          # BIAS_EINVAL <-> @check_args (src/common.jl:233-240)
    rc == 5 && throw(BoundsError())               # BIAS_ERANGE <-> the 10k Stirling table (src/HDP.jl:355)
    error("libbias_cuda error $rc: $msg")
end

function enabled()
    n = Cint[0]
    rc = try ccall((:bias_device_count, LIB), Cint, (Ptr{Cint},), n) catch; Cint(2) end
    rc == 0 && n[1] > 0
end

const CTX = Ptr{Void}[C_NULL]
function context(seed::Integer=0, device::Integer=0)
    if CTX[1] == C_NULL
        h = Ptr{Void}[C_NULL]
        check(ccall((:bias_ctx_create, LIB), Cint, (Cint, UInt64, Ptr{Void}, Ptr{Ptr{Void}}),
                    device, seed, C_NULL, h))
        CTX[1] = h[1]
    end
    CTX[1]
end

# ---- flattening of the reference's containers (1-based -> 0-based) -------------------------
flat_ptr(zz) = (p = zeros(Int64, length(zz) + 1); for j = 1:length(zz) p[j+1] = p[j] + length(zz[j]) end; p)
flat_zz(zz::Vector{Vector{Int}}) = Int32[z - 1 for g in zz for z in g]
flat_zz(zz::Vector{Int}) = Int32[z - 1 for z in zz]

conj_spec(c, datum) = isa(c, Main.BIAS.MultinomialDirichlet) ?
    Conjugate(BIAS_MD, datum, c.dd, c.aa, 0.0, 1.0, 1.0) :              # c.aa is already aa/dd (src/conjugates.jl:142)
    Conjugate(BIAS_G1G1, datum, 0, 0.0, c.m0, c.v0, c.vv)

# returns (Data, keepalive arrays)
function flat_data(items::Vector{Int})
    w = Int32[x - 1 for x in items]
    Data(length(w), pointer(w), C_NULL, C_NULL, C_NULL, C_NULL), (w,), BIAS_INT
end
function flat_data(items::Vector{Float64})
    Data(length(items), C_NULL, pointer(items), C_NULL, C_NULL, C_NULL), (items,), BIAS_F64
end
function flat_data(items::Vector{Main.BIAS.Sent})
    sp = zeros(Int64, length(items) + 1)
    for i = 1:length(items) sp[i+1] = sp[i] + length(items[i].w) end
    sw = Int32[w - 1 for s in items for w in s.w]
    sc = Int32[c for s in items for c in s.c]
    Data(length(items), C_NULL, C_NULL, pointer(sp), pointer(sw), pointer(sc)), (sp, sw, sc), BIAS_SENT
end

scatter!(zz::Vector{Int}, flat::Vector{Int32}) = (for i = 1:length(zz) zz[i] = flat[i] + 1 end)
function scatter!(zz::Vector{Vector{Int}}, flat::Vector{Int32})
    t = 0
    for g in zz, i = 1:length(g)
        t += 1; g[i] = flat[t] + 1
    end
end

# ---- collapsed_gibbs_sampler!(lda::LDA, xx, zz, n_burnins, n_lags, n_samples, ...)  src/LDA.jl:67-148
# With word-id observations, K <= 1024 and more than one GPU in BIAS_CUDA_DEVICES ("0,1,2,3,4,5,6,7"; default: every
# device the library sees) the documents are split over the GPUs inside the library (bias_lda_run_multi: one count
# exchange per sweep over NVLink peer memory); nothing else in the Julia program changes and no second process is needed.
function devices()
    if haskey(ENV, "BIAS_CUDA_DEVICES")
        return Int32[parse(Int32, t) for t in split(ENV["BIAS_CUDA_DEVICES"], ",")]
    end
    n = Cint[0]
    check(ccall((:bias_device_count, LIB), Cint, (Ptr{Cint},), n))
    Int32[i for i in 0:min(n[1], 8) - 1]
end

function lda!(lda, xx, zz, n_burnins::Int, n_lags::Int, n_samples::Int; seed=0)
    n_sweeps = n_burnins + n_samples * (n_lags + 1)                     # src/LDA.jl:82
    items = vcat(xx...)
    d, keep, datum = flat_data(items)
    c = conj_spec(lda.component, datum)
    ptr = flat_ptr(zz)
    z = flat_zz(zz)
    stats = Array(SweepStats, n_sweeps)
    devs = devices()
    if datum == BIAS_INT && lda.KK <= 1024 && length(devs) > 1
        check(ccall((:bias_lda_run_multi, LIB), Cint,
                    (Ptr{Int32}, Int32, UInt64, Ptr{Conjugate}, Ptr{Data}, Int64, Ptr{Int64}, Ptr{Int32}, Int32, Float64, Int32, Ptr{SweepStats}),
                    devs, length(devs), seed, [c], [d], length(zz), ptr, z, lda.KK, lda.aa, n_sweeps, stats))
    else
        check(ccall((:bias_lda_run, LIB), Cint,
                    (Ptr{Void}, Ptr{Conjugate}, Ptr{Data}, Int64, Ptr{Int64}, Ptr{Int32}, Int32, Float64, Int32, Ptr{SweepStats}),
                    context(seed), [c], [d], length(zz), ptr, z, lda.KK, lda.aa, n_sweeps, stats))
    end
    scatter!(zz, z)
    stats
end

# Where the most recent multi-GPU lda! spent its time: (set-up, sweeps, download and teardown) in milliseconds — the elapsed
# time the reference prints per iteration (src/LDA.jl:108-110), for the whole call.
function lda_timing()
    ms = zeros(Float64, 3)
    check(ccall((:bias_lda_run_multi_timing, LIB), Cint, (Ptr{Float64},), ms))
    (ms[1], ms[2], ms[3])
end

# ---- collapsed_gibbs_sampler!(bmm::BMM, xx, zz, ...)  src/BMM.jl:46-130
function bmm!(bmm, xx, zz::Vector{Int}, n_burnins::Int, n_lags::Int, n_samples::Int; seed=0)
    n_sweeps = n_burnins + n_samples * (n_lags + 1)
    d, keep, datum = flat_data(xx)
    c = conj_spec(bmm.component, datum)
    z = flat_zz(zz)
    stats = Array(SweepStats, n_sweeps)
    check(ccall((:bias_bmm_run, LIB), Cint,
                (Ptr{Void}, Ptr{Conjugate}, Ptr{Data}, Ptr{Int32}, Int32, Float64, Int32, Ptr{SweepStats}),
                context(seed), [c], [d], z, bmm.KK, bmm.aa, n_sweeps, stats))
    scatter!(zz, z)
    stats
end

# ---- collapsed_gibbs_sampler!(hdp::HDP, xx, zz, ...)  src/HDP.jl:166-399
# One call per kept sample reproduces the (KK_list, KK_dict, betas, gammas, alphas) return value.
function hdp!(hdp, xx, zz, n_burnins::Int, n_lags::Int, n_samples::Int,
              sample_hyperparam::Bool=true, n_internals::Int=10; Kmax::Int=256, seed=0)
    items = vcat(xx...)
    d, keep, datum = flat_data(items)
    c = conj_spec(hdp.component, datum)
    ptr = flat_ptr(zz)
    z = flat_zz(zz)
    hp = HdpParams(hdp.KK, Kmax, hdp.gg, hdp.g1, hdp.g2, hdp.aa, hdp.a1, hdp.a2, sample_hyperparam, n_internals)
    KK_list = zeros(Int, n_samples); KK_dict = Dict{Int, Vector{Vector{Int}}}()
    betas = Array(Vector{Float64}, n_samples); gammas = zeros(n_samples); alphas = zeros(n_samples)
    beta = zeros(Float64, Kmax + 1)
    run(n) = check(ccall((:bias_hdp_run, LIB), Cint,
                (Ptr{Void}, Ptr{Conjugate}, Ptr{Data}, Int64, Ptr{Int64}, Ptr{Int32}, Ptr{HdpParams}, Int32, Ptr{Float64}, Ptr{Void}),
                context(seed), [c], [d], length(zz), ptr, z, pointer_from_objref(hp), n, beta, C_NULL))
    run(n_burnins)
    for s = 1:n_samples
        run(n_lags + 1)
        scatter!(zz, z)
        KK_list[s] = hp.KK; KK_dict[hp.KK] = deepcopy(zz)
        betas[s] = beta[1:hp.KK+1]; gammas[s] = hp.gg; alphas[s] = hp.aa
    end
    scatter!(zz, z)
    hdp.KK, hdp.gg, hdp.aa = hp.KK, hp.gg, hp.aa
    KK_list, KK_dict, betas, gammas, alphas
end

# ---- CRF_gibbs_sampler!(hdp::HDP, xx, zz, ...)  src/HDP.jl:402-709
# Resident model created with BIAS_MODEL_CRF (= 4): one bias_model_sweep per iteration.
function crf!(hdp, xx, zz, n_burnins::Int, n_lags::Int, n_samples::Int,
              sample_hyperparam::Bool=true, n_internals::Int=10; Kmax::Int=256, seed=0)
    items = vcat(xx...)
    d, keep, datum = flat_data(items)
    c = conj_spec(hdp.component, datum)
    ptr = flat_ptr(zz)
    z = flat_zz(zz)
    hp = HdpParams(hdp.KK, Kmax, hdp.gg, hdp.g1, hdp.g2, hdp.aa, hdp.a1, hdp.a2, sample_hyperparam, n_internals)
    m = Ptr{Void}[C_NULL]
    check(ccall((:bias_model_create, LIB), Cint,
                (Ptr{Void}, Int32, Ptr{Conjugate}, Ptr{Data}, Int64, Ptr{Int64}, Ptr{Int32}, Int32, Float64, Ptr{HdpParams}, Ptr{Ptr{Void}}),
                context(seed), 4, [c], [d], length(zz), ptr, z, hdp.KK, hdp.aa, pointer_from_objref(hp), m))
    KK_list = zeros(Int, n_samples); KK_dict = Dict{Int, Vector{Vector{Int}}}()
    beta = zeros(Float64, Kmax + 1)
    try
        for iteration = 1:(n_burnins + n_samples * (n_lags + 1))
            check(ccall((:bias_model_sweep, LIB), Cint, (Ptr{Void}, Int32), m[1], 1))
            if (iteration - n_burnins) % (n_lags + 1) == 0 && iteration > n_burnins
                check(ccall((:bias_model_get_zz, LIB), Cint, (Ptr{Void}, Ptr{Int32}), m[1], z))
                check(ccall((:bias_model_get_hdp, LIB), Cint, (Ptr{Void}, Ptr{HdpParams}, Ptr{Float64}), m[1], pointer_from_objref(hp), beta))
                scatter!(zz, z)
                s = div(iteration - n_burnins, n_lags + 1)
                KK_list[s] = hp.KK; KK_dict[hp.KK] = deepcopy(zz)
            end
        end
        check(ccall((:bias_model_get_zz, LIB), Cint, (Ptr{Void}, Ptr{Int32}), m[1], z))
        check(ccall((:bias_model_get_hdp, LIB), Cint, (Ptr{Void}, Ptr{HdpParams}, Ptr{Float64}), m[1], pointer_from_objref(hp), beta))
        scatter!(zz, z)
        hdp.KK, hdp.gg, hdp.aa = hp.KK, hp.gg, hp.aa
    finally
        ccall((:bias_model_destroy, LIB), Cint, (Ptr{Void},), m[1])
    end
    KK_list, KK_dict
end

# ---- RCRP_gibbs_sampler!(rcrp::RCRP, xx, zz, ...)  src/RCRP.jl:77-416
# Resident model: upload once, one bias_model_sweep per iteration, zz read back at every kept sample.
function rcrp!(rcrp, xx, zz, n_burnins::Int, n_lags::Int, n_samples::Int,
               sample_hyperparam::Bool=true, n_internals::Int=10; Kmax::Int=256, seed=0)
    items = vcat(xx...)
    d, keep, datum = flat_data(items)
    c = conj_spec(rcrp.component, datum)
    ptr = flat_ptr(zz)
    z = flat_zz(zz)
    rp = RcrpParams(rcrp.KK, Kmax, rcrp.a1, rcrp.a2, sample_hyperparam, n_internals)
    aa = copy(rcrp.aa)
    m = Ptr{Void}[C_NULL]
    check(ccall((:bias_rcrp_create, LIB), Cint,
                (Ptr{Void}, Ptr{Conjugate}, Ptr{Data}, Int64, Ptr{Int64}, Ptr{Int32}, Ptr{RcrpParams}, Ptr{Float64}, Ptr{Ptr{Void}}),
                context(seed), [c], [d], length(zz), ptr, z, pointer_from_objref(rp), aa, m))
    KK_list = zeros(Int, n_samples); KK_dict = Dict{Int, Vector{Vector{Int}}}()
    try
        for iteration = 1:(n_burnins + n_samples * (n_lags + 1))                 # src/RCRP.jl:159
            if iteration > n_burnins                                            # :166-168
                check(ccall((:bias_model_set_sample_hyperparam, LIB), Cint, (Ptr{Void}, Int32), m[1], 0))
            end
            check(ccall((:bias_model_sweep, LIB), Cint, (Ptr{Void}, Int32), m[1], 1))
            if (iteration - n_burnins) % (n_lags + 1) == 0 && iteration > n_burnins
                check(ccall((:bias_model_get_zz, LIB), Cint, (Ptr{Void}, Ptr{Int32}), m[1], z))
                check(ccall((:bias_model_get_rcrp, LIB), Cint, (Ptr{Void}, Ptr{RcrpParams}, Ptr{Float64}), m[1], pointer_from_objref(rp), aa))
                scatter!(zz, z)
                s = div(iteration - n_burnins, n_lags + 1)
                KK_list[s] = rp.KK; KK_dict[rp.KK] = deepcopy(zz)
            end
        end
        check(ccall((:bias_model_get_zz, LIB), Cint, (Ptr{Void}, Ptr{Int32}), m[1], z))
        check(ccall((:bias_model_get_rcrp, LIB), Cint, (Ptr{Void}, Ptr{RcrpParams}, Ptr{Float64}), m[1], pointer_from_objref(rp), aa))
        scatter!(zz, z)
        rcrp.KK = rp.KK
        rcrp.aa[:] = aa
    finally
        ccall((:bias_model_destroy, LIB), Cint, (Ptr{Void},), m[1])
    end
    KK_list, KK_dict
end

# ---- RCRF_gibbs_sampler!(rcrf::RCRF, xx, zz, ...)  src/RCRF.jl:122-938
# xx::Vector{Vector{Vector{T}}} is flattened: the restaurants of all epochs are the groups, epoch_ptr their epochs.
function rcrf!(rcrf, xx, zz, n_burnins::Int, n_lags::Int, n_samples::Int, sample_hyperparam::Bool=true,
               n_internals::Int=10; join_tables::Bool=true, Kmax::Int=256, seed=0)
    groups = vcat(xx...); zgroups = vcat(zz...)
    epoch_ptr = Int64[0; cumsum([length(ep) for ep in xx])]
    d, keep, datum = flat_data(vcat(groups...))
    c = conj_spec(rcrf.component, datum)
    ptr = flat_ptr(zgroups)
    z = flat_zz(zgroups)
    rp = RcrfParams(rcrf.KK, Kmax, rcrf.g1, rcrf.g2, rcrf.a1, rcrf.a2, sample_hyperparam, n_internals, join_tables)
    gg = copy(rcrf.gg); aa = copy(rcrf.aa); KK = Int32[rcrf.KK]
    m = Ptr{Void}[C_NULL]
    check(ccall((:bias_rcrf_create, LIB), Cint,
                (Ptr{Void}, Ptr{Conjugate}, Ptr{Data}, Int64, Ptr{Int64}, Int64, Ptr{Int64}, Ptr{Int32}, Ptr{RcrfParams},
                 Ptr{Float64}, Ptr{Float64}, Ptr{Ptr{Void}}),
                context(seed), [c], [d], length(zgroups), ptr, length(xx), epoch_ptr, z, [rp], gg, aa, m))
    KK_list = zeros(Int, n_samples); KK_dict = Dict{Int, Vector{Vector{Vector{Int}}}}()
    try
        for iteration = 1:(n_burnins + n_samples * (n_lags + 1))
            if iteration > n_burnins                                            # src/RCRF.jl:226-228
                check(ccall((:bias_model_set_sample_hyperparam, LIB), Cint, (Ptr{Void}, Int32), m[1], 0))
            end
            check(ccall((:bias_model_sweep, LIB), Cint, (Ptr{Void}, Int32), m[1], 1))
            if (iteration - n_burnins) % (n_lags + 1) == 0 && iteration > n_burnins
                check(ccall((:bias_model_get_zz, LIB), Cint, (Ptr{Void}, Ptr{Int32}), m[1], z))
                check(ccall((:bias_model_get_rcrf, LIB), Cint, (Ptr{Void}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}), m[1], KK, gg, aa))
                scatter!(zgroups, z)            # zgroups aliases the inner vectors of zz
                s = div(iteration - n_burnins, n_lags + 1)
                KK_list[s] = KK[1]; KK_dict[Int(KK[1])] = deepcopy(zz)
            end
        end
        check(ccall((:bias_model_get_zz, LIB), Cint, (Ptr{Void}, Ptr{Int32}), m[1], z))
        check(ccall((:bias_model_get_rcrf, LIB), Cint, (Ptr{Void}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}), m[1], KK, gg, aa))
        scatter!(zgroups, z)
        rcrf.KK = KK[1]; rcrf.gg[:] = gg; rcrf.aa[:] = aa
    finally
        ccall((:bias_model_destroy, LIB), Cint, (Ptr{Void},), m[1])
    end
    KK_list, KK_dict
end

end # module
