// boost/thread.hpp — stand-in (TEST INFRASTRUCTURE, see ../Eigen/Core): the Boost.Thread / bind / function names that
// Include/IndexThreadReduce.h and Include/Frame.h use, mapped onto the C++11 standard library.
#pragma once
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
namespace boost {
using std::thread;
using std::mutex;
using std::unique_lock;
using std::condition_variable;
using std::function;
using std::bind;
}  // namespace boost
using namespace std::placeholders;
