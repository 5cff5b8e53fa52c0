// stand-in: tests/warp_test.cpp of the reference includes this file (mLib / Opt application scaffolding) but uses nothing from it
