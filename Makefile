# Top-level build, same target name as the reference's root Makefile (`make` -> gps_test).
#   make            libgpsacq.so (nvcc, sm_100a) + gps_test (C++ host)
#   make oracle     CPU oracle (test infrastructure; builds oracle/_ref only where /root/reference exists)
#   make emu        CPU replay of the kernel math (test infrastructure)
PKG := gnss-gps-sdr_b200

all: gps_test

lib:
	$(MAKE) -C $(PKG)/csrc

gps_test: lib
	$(MAKE) -C $(PKG)/c

oracle:
	$(MAKE) -C oracle

emu: tests/emu/libemu.so
tests/emu/libemu.so: tests/emu/emu.cpp $(wildcard $(PKG)/csrc/*.h)
	g++ -std=c++17 -O2 -fPIC -shared -I$(PKG)/csrc tests/emu/emu.cpp -o $@

clean:
	$(MAKE) -C $(PKG)/csrc clean
	$(MAKE) -C $(PKG)/c clean
	$(MAKE) -C oracle clean
	rm -f tests/emu/libemu.so gps_test.bin

.PHONY: all lib gps_test oracle emu clean
