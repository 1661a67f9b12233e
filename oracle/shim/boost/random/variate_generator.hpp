#ifndef PARM_ORACLE_BOOST_VARGEN
#define PARM_ORACLE_BOOST_VARGEN
#include <type_traits>
namespace boost {
template <class Engine, class Distribution>
class variate_generator {
    typedef typename std::remove_reference<Engine>::type engine_t;
    engine_t *_eng;
    Distribution _dist;
   public:
    typedef typename Distribution::result_type result_type;
    variate_generator(engine_t &e, Distribution d) : _eng(&e), _dist(d) {}
    result_type operator()() { return _dist(*_eng); }
    Distribution &distribution() { return _dist; }
    engine_t &engine() { return *_eng; }
};
}
#endif
