// TEST INFRASTRUCTURE: cereal/types/vector.hpp stand-in (handled inside the archive shim).
#pragma once
#include <vector>
