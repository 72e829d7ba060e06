// Stage 3 -- Boltzmann inversion epilogue, one thread per Bond.
//
// Reference: Bond.boltzmann_invert (pycgtool/bondset.py:62-91) with the statistic selection of
// bondset.py:172-179 (lengths: np.nanmean / np.nanvar; angles and dihedrals: circular mean /
// variance, pycgtool/util.py:24-53) and the force-constant formulas of
// pycgtool/functionalforms.py:92-148.  Inputs are the float64 partial sums produced by
// cgb_measure_fused / cgb_measure / cgb_circular_pass2 (already reduced across GPUs).

#include <math.h>

#include "cgb_common.cuh"

namespace cgb {

// True when the histogram row proves that no sample lies on the arc of angular length |e| that starts at the
// antipode `a0` of the shift and runs towards the antipode of the mean, widened by one bin on both sides.
static __device__ bool arc_is_empty(const double* __restrict__ row, double lo, double hi, int nbins, double n_valid,
                             double a0, double e) {
    const double pi = 3.141592653589793238462643383279, two_pi = 2.0 * pi;
    const double w = (hi - lo) / (double)nbins;
    if (!(w > 0.0)) return false;
    double total = 0.0;
    bool hit = false;
    const double centre = a0 + 0.5 * e, half = 0.5 * fabs(e) + 1.5 * w + 1e-5;
    for (int i = 0; i < nbins; ++i) {
        const double cnt = row[i];
        total += cnt;
        if (cnt != 0.0) {
            double d = (lo + ((double)i + 0.5) * w) - centre;
            d -= rint(d / two_pi) * two_pi;
            if (fabs(d) <= half) hit = true;
        }
    }
    return !hit && total == n_valid;   // every finite sample is in the histogram, none near the arc
}

__global__ void boltzmann_kernel(const double* __restrict__ partials, const float* __restrict__ shift,
                                 const int32_t* __restrict__ natoms, const int32_t* __restrict__ form,
                                 double temperature, int64_t n_bonds, const double* __restrict__ hist,
                                 const float* __restrict__ hist_lo, const float* __restrict__ hist_hi, int nbins,
                                 int pass2, double* __restrict__ out) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_bonds) return;
    const double* p = partials + b * CGB_NPART;
    double* o = out + b * CGB_NOUT;
    const double n = p[0];
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const double flag_in = o[4];
    if (!(n > 0.0)) {  // no finite value measured: the reference leaves eqm / fconst as None
        o[0] = o[1] = o[2] = o[3] = nan;
        o[4] = 0.0;
        return;
    }
    const int na = natoms[b];
    const double s = shift ? (double)shift[b] : 0.0;
    const double pi = 3.141592653589793238462643383279;
    double mean, var, flag = 0.0;
    if (na == 2) {
        const double md = p[1] / n;  // mean of (v - shift)
        mean = s + md;
        var = p[2] / n - md * md;
    } else {
        mean = atan2(p[4] / n, p[3] / n);  // util.py:33-37
        if (na == 3) {
            // v in [0, pi] and mean in [0, pi]: v - mean never wraps, so
            // <(v - mean)^2> = <d^2> - 2 e <d> + e^2 with d = v - shift, e = mean - shift.
            const double e = mean - s;
            var = (p[2] - 2.0 * e * p[1]) / n + e * e;
        } else if (pass2 == 2 || (pass2 == 1 && flag_in != 0.0)) {
            var = p[5] / n;  // pass 2 over the stored values (util.py:49-53)
            flag = 2.0;
        } else {
            // one pass: d = wrap(v - shift); exact unless a sample lies between the two antipodes
            double e = mean - s;
            e -= rint(e / (2.0 * pi)) * (2.0 * pi);
            var = (p[2] - 2.0 * e * p[1]) / n + e * e;
            bool exact = e == 0.0;
            if (!exact && hist)
                exact = arc_is_empty(hist + b * nbins, (double)hist_lo[b], (double)hist_hi[b], nbins, n, s - pi, e);
            flag = exact ? 0.0 : 1.0;
        }
    }
    if (var < 0.0) var = 0.0;  // rounding of a zero variance
    double eqm = mean;
    if (na == 4) {
        // bondset.py:81-84: eqm -= pi; eqm = ((eqm + pi) % 2pi) - pi, Python floor-modulo
        const double two_pi = 2.0 * pi;
        double t = (mean - pi) + pi;
        t = t - floor(t / two_pi) * two_pi;
        eqm = t - pi;
    }
    const double rt = 8.314 * temperature / 1000.0;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    double fc;
    switch (form[b]) {
        case CGB_FORM_HARMONIC:
            fc = var == 0.0 ? inf : rt / var;  // bondset.py:89-91
            break;
        case CGB_FORM_COSHARMONIC: {
            const double sn = sin(mean);
            const double den = sn * sn * var;
            fc = den == 0.0 ? inf : rt / den;
            break;
        }
        case CGB_FORM_MARTINI_LENGTH:
            fc = 1250.0;
            break;
        case CGB_FORM_MARTINI_ANGLE:
            fc = 25.0;
            break;
        case CGB_FORM_MARTINI_DIHEDRAL:
            fc = 50.0;
            break;
        default:
            fc = nan;  // user-defined form: the host evaluates it from mean / variance
    }
    o[0] = eqm;
    o[1] = fc;
    o[2] = mean;
    o[3] = var;
    o[4] = flag;
}

}  // namespace cgb

extern "C" int cgb_boltzmann(const double* partials, const float* shift, const int32_t* natoms,
                             const int32_t* form, double temperature, int64_t n_bonds, const double* hist,
                             const float* hist_lo, const float* hist_hi, int32_t n_bins, int pass2, double* out,
                             void* stream) {
    if (n_bonds < 0 || pass2 < 0 || pass2 > 2) return CGB_EINVAL;
    if (n_bonds == 0) return CGB_OK;
    if (!partials || !natoms || !form || !out) return CGB_EINVAL;
    if (hist && (!hist_lo || !hist_hi || n_bins < 1)) return CGB_EINVAL;
    const int threads = 64;
    const unsigned blocks = (unsigned)((n_bonds + threads - 1) / threads);
    cgb::boltzmann_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(
        partials, shift, natoms, form, temperature, n_bonds, hist, hist_lo, hist_hi, n_bins, pass2, out);
    return (int)cudaGetLastError();
}

extern "C" int cgb_abi_version(void) { return CGB_ABI_VERSION; }

extern "C" const char* cgb_strerror(int code) {
    switch (code) {
        case CGB_OK:
            return "success";
        case CGB_EINVAL:
            return "invalid argument (NULL pointer, negative size or inconsistent table)";
        case CGB_ERANGE:
            return "index outside the atom / bead range";
        case CGB_ETOOBIG:
            return "bead or tile too large for the shared-memory tile";
        case CGB_EALIGN:
            return "pointer not aligned to its element type";
        case CGB_ENCCL:
            return "NCCL is unavailable (libnccl.so.2 not loadable) or an NCCL call failed";
        default:
            break;
    }
    if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
    return "unknown error";
}
