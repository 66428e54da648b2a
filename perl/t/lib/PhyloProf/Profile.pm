package PhyloProf::Profile;

# TEST STUB (perl/t only): the part of the PhyloProf::Profile interface that the drop-ins touch -- new(-raw/-yes/-no),
# new(-profile), new(-file [, -rank, -taxonomy]), get_hash / get_yes / get_total / is_yes (lib/PhyloProf/Profile.pm:93-203,
# 284-417, 433-549) -- so that they can be exercised on the GPU box, where the reference tree does not exist.  In production
# the reference's own class is used.
use strict;
use warnings;

sub new {
    my ( $class, %a ) = @_;
    my $self = bless { _yes => 0, _no => 0 }, $class;
    if ( defined $a{-raw} ) {
        @$self{qw(_prof _yes _no)} = ( $a{-raw}, $a{-yes}, $a{-no} );
        return $self;
    }
    my @items;
    if ( defined $a{-file} ) {
        open( my $fh, "<", $a{-file} ) or die "Can' open $a{-file}\n";
        while ( my $line = <$fh> ) {
            chomp $line;
            next if $line =~ /^\#/;
            my @f = split( /\t/, $line );
            die "$a{-file} must contain only two fields\n" if @f != 2;
            push @items, [@f];
        }
        close($fh);
    } else {
        @items = map { [ $_, $a{-profile}{$_} ] } keys %{ $a{-profile} };
    }
    my %h;
    foreach my $it (@items) {
        my ( $t, $v ) = @$it;
        if ( $a{-rank} ) { $t = $a{-taxonomy}->collapse_taxon( $t, $a{-rank} ) || $a{-taxonomy}->collapse_taxon( $t, "genus" ) || $t }
        next unless $t;
        next if exists $h{$t} && $h{$t} != 0;
        $h{$t} = $v;
    }
    $self->{_prof} = \%h;
    $self->{_yes}  = grep { $_ != 0 } values %h;
    $self->{_no}   = ( scalar keys %h ) - $self->{_yes};
    return $self;
}
sub get_hash  { $_[0]{_prof} }
sub get_yes   { $_[0]{_yes} }
sub get_total { $_[0]{_yes} + $_[0]{_no} }
sub is_yes    { my ( $s, $t ) = @_; return ( defined $t && exists $s->{_prof}{$t} && $s->{_prof}{$t} != 0 ) ? 1 : 0 }
1;
