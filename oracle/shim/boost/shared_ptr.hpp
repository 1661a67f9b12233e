// Boost is a third-party dependency of the reference and is absent here; the
// reference only uses shared_ptr and pointer casts, which std provides.
// TEST INFRASTRUCTURE ONLY (oracle build).
#ifndef PARM_ORACLE_BOOST_SHARED_PTR
#define PARM_ORACLE_BOOST_SHARED_PTR
#include <memory>
namespace boost {
using std::shared_ptr;
using std::static_pointer_cast;
using std::dynamic_pointer_cast;
using std::const_pointer_cast;
}
#endif
