// cvcompat: see opencv2/core/core.hpp in this directory tree
#pragma once
#include <opencv2/core/core.hpp>
