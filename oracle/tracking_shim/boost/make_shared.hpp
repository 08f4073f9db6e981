// stand-in for <boost/make_shared.hpp> (include/Tracking.h:42): the tracker only holds and forwards a boost::shared_ptr<PointCloudMapping>.  TEST INFRASTRUCTURE.
#pragma once
#include <memory>
namespace boost { template <class T> using shared_ptr = std::shared_ptr<T>; using std::make_shared; }
