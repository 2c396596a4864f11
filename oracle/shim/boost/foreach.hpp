// BOOST_FOREACH stub (TEST INFRASTRUCTURE)
#pragma once
#define BOOST_FOREACH(decl, expr) for (decl : expr)
