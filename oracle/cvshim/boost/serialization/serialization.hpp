// ORACLE: inert stand-in for the two boost::serialization names the vendored DBoW2 headers mention (BowVector.h:62-66,
// FeatureVector.h:27-31); their serialize() templates are never instantiated here.
#pragma once
namespace boost { namespace serialization {
class access {};
template <class Base, class Derived> Base& base_object(Derived& d) { return d; }
}}  // namespace boost::serialization
