package PhyloProf::Profile;

# TEST STUB (perl/t only): the minimal PhyloProf::Profile interface that PhyloProf::Algorithm::ppp(_b200) touches
# -- get_hash / get_yes / get_total (lib/PhyloProf/Profile.pm:342-417) -- so that the drop-in can be exercised on the
# GPU box, where the reference tree does not exist.  In production the reference's own class is used.
use strict;
use warnings;

sub new {
    my ( $class, %a ) = @_;
    my $self = bless { _yes => 0, _no => 0 }, $class;
    if ( defined $a{-raw} ) {
        @$self{qw(_prof _yes _no)} = ( $a{-raw}, $a{-yes}, $a{-no} );
    } else {
        my %h = %{ $a{-profile} };
        $self->{_prof} = \%h;
        $self->{_yes}  = grep { $_ != 0 } values %h;
        $self->{_no}   = ( scalar keys %h ) - $self->{_yes};
    }
    return $self;
}
sub get_hash  { $_[0]{_prof} }
sub get_yes   { $_[0]{_yes} }
sub get_total { $_[0]{_yes} + $_[0]{_no} }
1;
