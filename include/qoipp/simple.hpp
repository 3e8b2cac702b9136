// <qoipp/simple.hpp> -- kept so that code written against mrizaln/qoipp's header layout compiles unchanged;
// the whole API lives in <qoipp/qoipp.hpp>.
#pragma once
#include "qoipp/qoipp.hpp"
