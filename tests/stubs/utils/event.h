// tests/stubs/utils/event.h -- TEST INFRASTRUCTURE: the wrappers include it and use nothing from it
#pragma once
