// Stand-in for <kdl/chain.hpp> (orocos_kdl is not in this container), written for the oracle build only.
// distance_heuristics.cpp:159-214 asks a KDL::Chain only for the joint-type NAME of each segment, so this models
// exactly that: a list of segments whose joints answer getTypeName() with KDL's strings ("None", "RotAxis",
// "TransAxis", ... — orocos_kdl joint.hpp Joint::getTypeName()).
#pragma once
#include <string>
#include <vector>
namespace KDL {
class Joint {
 public:
  explicit Joint(const std::string &type_name = "None") : type_name_(type_name) {}
  const std::string getTypeName() const { return type_name_; }
 private:
  std::string type_name_;
};
class Segment {
 public:
  explicit Segment(const Joint &j = Joint()) : joint_(j) {}
  const Joint &getJoint() const { return joint_; }
 private:
  Joint joint_;
};
class Chain {
 public:
  void addSegment(const Segment &s) { segments_.push_back(s); }
  unsigned int getNrOfSegments() const { return (unsigned int)segments_.size(); }
  const Segment &getSegment(unsigned int i) const { return segments_[i]; }
 private:
  std::vector<Segment> segments_;
};
}  // namespace KDL
