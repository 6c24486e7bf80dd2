// stand-in (TEST INFRASTRUCTURE)
#include <boost/thread.hpp>
