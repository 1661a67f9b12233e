#ifndef PARM_ORACLE_BOOST_ARRAY
#define PARM_ORACLE_BOOST_ARRAY
#include <array>
namespace boost { template <typename T, std::size_t N> using array = std::array<T, N>; }
#endif
