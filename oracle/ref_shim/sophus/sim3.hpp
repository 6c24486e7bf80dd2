// stand-in: see se3.hpp
#pragma once
#include "se3.hpp"
namespace Sophus { class Sim3d {}; }
