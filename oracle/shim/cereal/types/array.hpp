// TEST INFRASTRUCTURE: cereal/types/array.hpp stand-in (handled inside the archive shim).
#pragma once
#include <array>
