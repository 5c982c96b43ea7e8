#pragma once
#include "glew.h"
