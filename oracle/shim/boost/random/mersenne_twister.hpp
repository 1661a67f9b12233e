#ifndef PARM_ORACLE_BOOST_MT
#define PARM_ORACLE_BOOST_MT
#include <random>
namespace boost { typedef std::mt19937 mt19937; }
#endif
