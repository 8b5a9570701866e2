// special_functions.cuh -- modified Bessel functions K_n and exponential integrals E_n
// in FP64, for (a) filling the tabulated grids the reference builds with GSL
// (FSSW::initialize_special_function_arrays, FSSW.cpp:1609-1643) and (b) the
// out-of-table fall-back of FSSW::get_special_function_K1/K2/K3/En
// (FSSW.cpp:1646-1727).  GSL is a third-party dependency that is not part of
// the reference tree; these are restatements of the published algorithms
// (ascending series for x <= 2 and Steed's continued fraction CF2 for x > 2,
// Temme 1975 / Thompson & Barnett 1987; E_n by power series for x <= 1 and the
// modified-Lentz continued fraction otherwise).  Agreement with GSL is ~1e-14,
// far inside the 1e-6 yield tolerance.
#ifndef ISS_SPECIAL_FUNCTIONS_CUH_
#define ISS_SPECIAL_FUNCTIONS_CUH_

namespace iss {

// K_0(x) and K_1(x), x > 0.
__host__ __device__ inline void bessel_k01(double x, double &k0, double &k1) {
    const double euler = 0.57721566490153286061;
    if (x <= 2.0) {
        // ascending series
        const double y = 0.25*x*x;
        const double lnhx = log(0.5*x);
        // I0, I1 series and the K sums
        double term0 = 1.0;         // y^k/(k!)^2
        double hk = 0.0;            // harmonic number H_k
        double i0 = 1.0, s0 = 0.0;
        double term1 = 1.0;         // y^k/(k!(k+1)!)
        double i1 = 1.0;
        double s1 = 1.0 - 2.0*euler;   // k=0: psi(1)+psi(2) = -2 gamma + 1
        for (int k = 1; k < 40; k++) {
            term0 *= y/(static_cast<double>(k)*k);
            hk += 1.0/k;
            i0 += term0;
            s0 += term0*hk;
            term1 *= y/(static_cast<double>(k)*(k + 1));
            i1 += term1;
            // psi(k+1) + psi(k+2) = -2 gamma + 2 H_k + 1/(k+1)
            s1 += term1*(-2.0*euler + 2.0*hk + 1.0/(k + 1));
            if (term0 < 1e-18*i0) break;
        }
        k0 = -(lnhx + euler)*i0 + s0;
        const double I1 = 0.5*x*i1;
        k1 = 1.0/x + lnhx*I1 - 0.25*x*s1;
        return;
    }
    // Steed's CF2 for mu = 0
    double b = 2.0*(1.0 + x);
    double d = 1.0/b;
    double h = d, delh = d;
    double q1 = 0.0, q2 = 1.0;
    const double a1 = 0.25;
    double q = a1, c = a1, a = -a1;
    double s = 1.0 + q*delh;
    for (int i = 2; i < 10000; i++) {
        a -= 2*(i - 1);
        c = -a*c/i;
        const double qnew = (q1 - b*q2)/a;
        q1 = q2;
        q2 = qnew;
        q += c*qnew;
        b += 2.0;
        d = 1.0/(b + a*d);
        delh = (b*d - 1.0)*delh;
        h += delh;
        const double dels = q*delh;
        s += dels;
        if (fabs(dels/s) < 1e-17) break;
    }
    h = a1*h;
    k0 = sqrt(3.14159265358979323846/(2.0*x))*exp(-x)/s;
    k1 = k0*(x + 0.5 - h)/x;
}

// K_1, K_2, K_3 by upward recurrence K_{n+1} = K_{n-1} + (2n/x) K_n (stable for K).
__host__ __device__ inline void bessel_k123(double x, double &k1, double &k2, double &k3) {
    double k0;
    bessel_k01(x, k0, k1);
    k2 = k0 + 2.0/x*k1;
    k3 = k1 + 4.0/x*k2;
}

// E_n(x), n >= 2, x > 0
__host__ __device__ inline double expint_en(int n, double x) {
    const double euler = 0.57721566490153286061;
    const int nm1 = n - 1;
    if (x > 1.0) {
        double b = x + n, c = 1e300, d = 1.0/b, h = d;
        for (int i = 1; i < 10000; i++) {
            const double a = -1.0*i*(nm1 + i);
            b += 2.0;
            d = 1.0/(a*d + b);
            c = b + a/c;
            const double del = c*d;
            h *= del;
            if (fabs(del - 1.0) < 1e-16) break;
        }
        return h*exp(-x);
    }
    double ans = 1.0/nm1;
    double fact = 1.0;
    for (int i = 1; i < 10000; i++) {
        fact *= -x/i;
        double del;
        if (i != nm1) {
            del = -fact/(i - nm1);
        } else {
            double psi = -euler;
            for (int ii = 1; ii <= nm1; ii++) psi += 1.0/ii;
            del = fact*(-log(x) + psi);
        }
        ans += del;
        if (fabs(del) < fabs(ans)*1e-16) break;
    }
    return ans;
}

}  // namespace iss
#endif  // ISS_SPECIAL_FUNCTIONS_CUH_
