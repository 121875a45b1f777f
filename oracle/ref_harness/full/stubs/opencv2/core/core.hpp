// stand-in for <opencv2/core/core.hpp> (TEST INFRASTRUCTURE, oracle/_ref): rv/Laserscan.h includes it but uses nothing of it
#pragma once
typedef unsigned char uchar;
