#pragma once  // threadIdx & co. come from emu.h
