// boost::normal_distribution's algorithm is Boost-version dependent
// (Box-Muller before 1.56, ziggurat after): the Gaussian stream of the
// reference is therefore "parity unpinned". This stand-in is Box-Muller in the
// pre-1.56 Boost formulation (cached second value), over uniform_01 draws.
#ifndef PARM_ORACLE_BOOST_NORMAL
#define PARM_ORACLE_BOOST_NORMAL
#include <cmath>
#include <cstddef>
// Noise injection hook (defined in oracle/ref_harness.cpp): when set, draws are
// taken from this array in draw order so the GPU path and the reference can be
// fed bit-identical Gaussians.
extern "C" const double *parm_oracle_noise;
extern "C" size_t parm_oracle_noise_len;
extern "C" size_t parm_oracle_noise_pos;
namespace boost {
template <class RealType = double>
class normal_distribution {
    RealType _mean, _sigma, _r1, _r2, _cached_rho;
    bool _valid;
   public:
    typedef RealType input_type;
    typedef RealType result_type;
    explicit normal_distribution(const RealType &mean = RealType(0), const RealType &sigma = RealType(1))
        : _mean(mean), _sigma(sigma), _r1(0), _r2(0), _cached_rho(0), _valid(false) {}
    RealType mean() const { return _mean; }
    RealType sigma() const { return _sigma; }
    void reset() { _valid = false; }
    template <class Engine>
    result_type operator()(Engine &eng) {
        using std::sqrt; using std::log; using std::sin; using std::cos;
        if (parm_oracle_noise && parm_oracle_noise_pos < parm_oracle_noise_len)
            return RealType(parm_oracle_noise[parm_oracle_noise_pos++]) * _sigma + _mean;
        if (!_valid) {
            _r1 = RealType(eng()) / RealType(4294967296.0);
            _r2 = RealType(eng()) / RealType(4294967296.0);
            _cached_rho = sqrt(-RealType(2) * log(RealType(1) - _r2));
            _valid = true;
        } else {
            _valid = false;
        }
        const RealType pi = RealType(3.14159265358979323846);
        return _cached_rho * (_valid ? cos(RealType(2) * pi * _r1) : sin(RealType(2) * pi * _r1)) * _sigma + _mean;
    }
};
template <class RealType = double, class R2 = RealType>
class uniform_01 {
   public:
    typedef RealType input_type;
    typedef RealType result_type;
    template <class Engine>
    result_type operator()(Engine &eng) { return RealType(eng()) / RealType(4294967296.0); }
};
}
#endif
