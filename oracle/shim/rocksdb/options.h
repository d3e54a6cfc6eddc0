#pragma once
#include "rocksdb/db.h"
